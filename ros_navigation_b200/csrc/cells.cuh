/*
 * cells.cuh -- how a grid_map layer lives in HBM.
 *
 * The reference keeps every layer as a column-major float matrix (grid_map_core/include/grid_map_core/GridMap.hpp:
 * 493-516, Eigen::MatrixXf).  The HIMM update only ever produces values of the closed set {NaN, 0, 10, ..., 180}
 * (clearCell / markCell, move_control/include/move_control/map_updater.h:49-71 map the set into itself), so a layer is
 * kept in one of two device formats:
 *
 *   CODED  (default)  one byte per cell, tile-major.  The grid is cut into 64 x 64 tiles; tile (tr, tc) of a robot is
 *                     a contiguous 4352-byte record (64 columns x 68-byte pitch), records ordered column of tiles
 *                     after column of tiles:  offset = ((tc * tiles_r + tr) * 4352) + (c & 63) * 68 + (r & 63).
 *                     The record is exactly the shared-memory image the HIMM tile kernel works on, so staging a tile
 *                     is ONE bulk async copy (TMA engine) in and one out, and a cell costs 1 byte of HBM traffic per
 *                     direction instead of 4.  Pad bytes and cells outside the grid hold code 1 and are never read as
 *                     cells.
 *   FLOAT             the reference's own layout ([robot][col][row] float).  A layer switches to it (for good) when
 *                     the host uploads a value outside the set or asks for the raw device pointer; the HIMM kernel
 *                     then converts on load / store and processes tiles with foreign values in place.
 *
 * Codes: 0 = NaN (unknown), k = value/10 + 1 for value in {0, 10, ..., 180} (1..19).  With this numbering
 *   clearCell(c) = max(c - 1, 1)                              (NaN -> 0; 0 stays 0)
 *   markCell(c)  = c <= 1 ? 4 : (c <= 16 ? c + 3 : c)         (NaN or 0 -> 30; <= 150 -> +30)
 */
#ifndef B200NAV_CELLS_CUH
#define B200NAV_CELLS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

namespace b200nav {

#define HIMM_TILE 64          /* tile edge in cells (one warp owns one tile)            */
#define HIMM_TILE_PITCH 68    /* bytes per tile column: 68/4 is odd -> conflict-free walks along rows and columns */
#define HIMM_TILE_BYTES (HIMM_TILE * HIMM_TILE_PITCH)
#define HIMM_CODE_NAN 0
#define HIMM_CODE_FREE 1

/* Small non-negative integers <-> float without I2F/F2I: float(0x4B000000 + k) == 8388608 + k exactly. */
__device__ __forceinline__ float small_int_to_float(int k) { return __int_as_float(0x4B000000 + k) - 8388608.0f; }

/* float -> code; returns 255 for a value outside the HIMM set */
__device__ __forceinline__ unsigned himm_encode(float v) {
  /* c = round(v/10) through the magic add; garbage for NaN / negative / huge inputs is rejected by the checks */
  const int c = __float_as_int(v * 0.1f + 8388608.0f) - 0x4B000000;
  const bool in_set = (unsigned)c <= 18u && small_int_to_float(c * 10) == v && __float_as_uint(v) != 0x80000000u;
  /* -0.0f compares equal to 0 but has another bit pattern: it stays out of the set so that it round-trips */
  return (v != v) ? (unsigned)HIMM_CODE_NAN : (in_set ? (unsigned)(c + 1) : 255u);
}
__device__ __forceinline__ float himm_decode(unsigned c) {
  return c == HIMM_CODE_NAN ? __int_as_float(0x7fc00000) : small_int_to_float((int)(c * 10u) - 10);
}

__host__ __device__ __forceinline__ size_t coded_cell_offset(int tiles_r, int r, int c) {
  return ((size_t)((c >> 6) * tiles_r + (r >> 6))) * HIMM_TILE_BYTES + (size_t)((c & 63) * HIMM_TILE_PITCH + (r & 63));
}

/* Read handle of one robot's layer in either format. */
struct LayerRef {
  const void* base; /* this robot's first byte */
  int coded;
  int rows;    /* FLOAT: column pitch */
  int tiles_r; /* CODED: tiles per column of tiles */
  __device__ __forceinline__ float at(int r, int c) const {
    if (coded) return himm_decode(static_cast<const uint8_t*>(base)[coded_cell_offset(tiles_r, r, c)]);
    return static_cast<const float*>(base)[(size_t)c * rows + r];
  }
};

/* ---- format conversion / maintenance (not hot) --------------------------------------------------------------- */

/* All tiles of `n_tiles_total` records: in-grid cells = NaN, everything else = free code. */
__global__ void coded_fill_kernel(uint8_t* __restrict__ p, size_t n_records, int rows, int cols, int tiles_r,
                                  int tiles_per_robot) {
  const size_t n = n_records * HIMM_TILE_BYTES;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int tile = (int)((i / HIMM_TILE_BYTES) % tiles_per_robot);
    const int in = (int)(i % HIMM_TILE_BYTES);
    const int r = (tile % tiles_r) * HIMM_TILE + in % HIMM_TILE_PITCH, c = (tile / tiles_r) * HIMM_TILE + in / HIMM_TILE_PITCH;
    const bool cell = (in % HIMM_TILE_PITCH) < HIMM_TILE && r < rows && c < cols;
    p[i] = cell ? HIMM_CODE_NAN : HIMM_CODE_FREE;
  }
}

/* rows [r0, r0+nr) x cols [c0, c0+nc) of one robot's coded layer := code.  grid = (ceil(nr/128), min(nc, 65535)). */
__global__ void coded_fill_rect_kernel(uint8_t* __restrict__ base, int tiles_r, int r0, int nr, int c0, int nc,
                                       unsigned code) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nr) return;
  for (int c = blockIdx.y; c < nc; c += gridDim.y) base[coded_cell_offset(tiles_r, r0 + r, c0 + c)] = (uint8_t)code;
}

/* n_robots column-major float matrices -> coded records; *bad is set when a value has no code (run with check_only
 * first: a destination that received a 255 must not be used). */
__global__ void coded_from_float_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, int rows, int cols,
                                        int tiles_r, size_t robot_bytes, size_t n_cells_total, int check_only,
                                        int* __restrict__ bad) {
  const size_t per = (size_t)rows * cols;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  bool any_bad = false;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_cells_total; i += stride) {
    const size_t robot = i / per, lin = i - robot * per;
    const int r = (int)(lin % rows), c = (int)(lin / rows);
    const unsigned code = himm_encode(src[i]);
    any_bad |= code == 255u;
    if (!check_only) dst[robot * robot_bytes + coded_cell_offset(tiles_r, r, c)] = (uint8_t)code;
  }
  if (any_bad) *bad = 1;
}

__global__ void coded_to_float_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, int rows, int cols,
                                      int tiles_r, size_t robot_bytes, size_t n_cells_total) {
  const size_t per = (size_t)rows * cols;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_cells_total; i += stride) {
    const size_t robot = i / per, lin = i - robot * per;
    const int r = (int)(lin % rows), c = (int)(lin / rows);
    dst[i] = himm_decode(src[robot * robot_bytes + coded_cell_offset(tiles_r, r, c)]);
  }
}

}  // namespace b200nav
#endif

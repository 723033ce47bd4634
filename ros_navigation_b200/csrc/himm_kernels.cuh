/*
 * himm_kernels.cuh -- HIMM certainty-grid update on sm_100a.
 *
 * Replaces (behaviour, not code): MapUpdater::lineOnMap / clearCell / markCell
 * (move_control/include/move_control/map_updater.h:38-71), the in-order sample loop of
 * LaserMapUpdater::updateMap (move_control/src/laser_map_updater.cpp:15-20) and grid_map::LineIterator
 * (grid_map_core/src/iterators/LineIterator.cpp:16-150).
 *
 * Design:
 *   K0 himm_prep_kernel        one thread per RangeSample (or cloud point, or raw scan reading): fp64 clip of both ray
 *                              ends into the map and position->index, producing a 24-byte BeamSeg (integer Bresenham
 *                              end points + mark cell); then every 64 x 64 tile the line touches gets the beam's bit
 *                              in the tile's beam mask, and every tile touched for the first time is appended to the
 *                              work list.
 *   K1 himm_tile_coded_kernel  persistent one-warp CTAs pull (robot, tile) work items.  The warp OWNS the tile: the
 *                              byte-coded tile record (cells.cuh) is staged in shared memory by one bulk async copy,
 *                              every beam of the tile's mask is applied strictly in sample order (clear along the
 *                              Bresenham cells, then the +30 mark), 32 beams at a time in lock step over the step
 *                              index, and the record is copied back.  Because a cell is only ever touched by its owner
 *                              warp, in order, no atomics are needed and the saturating clear / mark sequence is
 *                              reproduced bit-exactly - the result cannot depend on scheduling.
 *      himm_tile_kernel        the same walk for float layers (float4 staging with encode / decode; tiles holding
 *                              values outside the HIMM set are processed in place).
 */
#ifndef B200NAV_HIMM_KERNELS_CUH
#define B200NAV_HIMM_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200nav.h"
#include "cells.cuh"
#include "geometry.h"
#include "scan_project.h"

namespace b200nav {

#define HIMM_CHUNK 2048     /* max beams per chunk: a tile's beam set is a bit mask of <= 2048 bits          */
#define HIMM_MASK_WORDS (HIMM_CHUNK / 32)

struct HimmArgs {
  GridDims dims;
  const RobotGeom* geom;          /* [n_robots]                                   */
  int coded;                      /* layer format (cells.cuh)                      */
  void* layer;                    /* FLOAT: float [n_robots][cols][rows]; CODED: bytes [n_robots][tiles][4352] (cells.cuh) */
  const b200nav_sample* samples;  /* device; NULL when the cloud form below is used */
  const double* origins;          /* cloud form: [n_active][2] laser origin per robot  */
  const float2* xy;               /* cloud form: [total] float32 end points            */
  const uint8_t* clear_end;       /* cloud form: [total] ifClearEnd flags or NULL      */
  /* scan form (scan_project.h): raw ranges + sensor pose per robot; the binning kernel projects them itself */
  const float* scan_ranges;       /* [n_active][scan.n_ranges] or NULL                 */
  const double* scan_poses;       /* [n_active][3] sensor x, y, yaw in the map frame   */
  const int32_t* scan_sel;        /* [scan.n_used] selected range indices or NULL (identity) */
  ScanModel scan;
  const int32_t* offsets;         /* device [n_robots+1], or NULL in single mode  */
  BeamSeg* segs;                  /* device scratch [total]                       */
  /* binning scratch, all-zero between updates (the tile kernel clears what it consumes):
   *   beam_masks[robot][chunk][tile][HIMM_MASK_WORDS]  bit b = beam (chunk*HIMM_CHUNK + b) of the robot touches tile
   */
  uint32_t* beam_masks;
  int* error_flag;                /* set when a robot has more samples than n_chunks * HIMM_CHUNK              */
  /* work list of touched (robot, tile) pairs, filled by the prep kernel, consumed by the persistent tile kernel:
   *   touched[robot*n_tiles + tile]  0/1 first-touch flag (cleared by the consumer)
   *   worklist[]                     keys robot*n_tiles + tile, in first-touch order
   *   counters[0..3]                 heavy entries (front) / next entry to hand out / warps finished / light
   *                                  entries (appended from the back)                                           */
  uint32_t* touched;
  int* worklist;
  int worklist_cap;               /* entries in worklist[] (n_active * n_tiles)                              */
  int* counters;
  /* free_cols[robot*n_tiles + tile]: 64 bits of "known to hold exactly 0 (free)" per 64 x 64 tile.  FLOAT layers: bit c
   * = column c of the tile; CODED layers: bit bc * 8 + br = the 8 x 8 cells of block column bc / block row br
   * (FreeBlocks below).  All ones = the whole tile is free in both.  Maintained by the tile kernel at write-back,
   * reset by every other writer of the layer and whenever the layer changes its format.  Beams that only re-clear
   * known-free cells and carry no mark cannot change anything (clearing 0 gives 0) and are skipped. */
  unsigned long long* free_cols;
  int robot0;                     /* first robot handled by blockIdx.y == 0       */
  int n_active;                   /* robots handled by this launch                */
  int single_n;                   /* >= 0: single-robot mode, samples [0, n)      */
  int total;                      /* total samples                                */
  /* the prep kernel may be launched per group of robots (pipelined with the host->device copies):
   * it then handles robots [rel_lo, rel_hi) of this update (their beams are [beam_lo, beam_hi)) */
  int beam_lo, beam_hi, rel_lo, rel_hi;
  int max_per_robot;              /* upper bound of the beams of one robot (sizes the prep grid) */
  int tiles_r, tiles_c;           /* tiles per grid                               */
  int n_chunks;                   /* chunks per robot                             */
  int chunk_beams;                /* beams per chunk: multiple of 32, <= HIMM_CHUNK */
  int mask_words;                 /* chunk_beams / 32                              */
  int mw_all;                     /* small fleets: himm_tile_coded_mw_kernel takes EVERY item (no one-warp launch) */
  int defer_first_touch;          /* prep kernel: first-touch probes after the binning loop (small / medium fleets) */
};

/* ---------------------------------------------------------------------------------------------------------------
 * K0: RangeSample -> BeamSeg, and binning: every tile the Bresenham line (or the mark) touches gets the beam's bit
 * in its beam mask (RED.OR).  Reading a mask in ascending bit order later yields the tile's beams in sample order
 * without any sort.
 * Warp-aggregated atomics: the 32 beams of a warp are consecutive samples = the 32 bits of ONE mask word per tile.
 * The warp walks its beams' tile lists in lock step; match.any groups the lanes that emit the same tile in an
 * iteration and the lowest lane of each group issues one RED.OR with the group's lane mask.  ~10x fewer L2
 * reductions than one per (beam, tile).
 * ------------------------------------------------------------------------------------------------------------- */
/* First touch of (robot, tile) `rt` in this update: append it to the work list (heavy items from the front). */
__device__ __forceinline__ void himm_first_touch(const HimmArgs& a, int rt, bool heavy) {
  if (a.touched[rt] == 0u && atomicExch(&a.touched[rt], 1u) == 0u) {
    if (heavy) a.worklist[atomicAdd(&a.counters[0], 1)] = rt;
    else a.worklist[a.worklist_cap - 1 - atomicAdd(&a.counters[3], 1)] = rt;
  }
}

#ifndef HIMM_PREP_BLOCKS
#define HIMM_PREP_BLOCKS 12
#endif
#define HIMM_PREP_PROBE_WORDS 512 /* per-CTA "tile already probed" bits: grids of up to 16384 tiles */
__global__ void __launch_bounds__(128, HIMM_PREP_BLOCKS) himm_prep_kernel(HimmArgs a) {
  /* grid = (blocks per robot, robots of this launch): the robot is the block's y index - no search through the
   * offsets - and a warp never straddles two robots */
  const int lane = threadIdx.x & 31;
  const int rel = a.rel_lo + (int)blockIdx.y + (int)gridDim.y * (int)blockIdx.z;
  if (rel >= a.rel_hi) return;
  int beg = 0, end = a.single_n;
  if (a.single_n < 0) {
    beg = __ldg(&a.offsets[rel]);
    end = __ldg(&a.offsets[rel + 1]);
  }
  const int i = beg + blockIdx.x * blockDim.x + threadIdx.x;
  if (blockIdx.x == 0 && threadIdx.x == 0 && end - beg > a.n_chunks * a.chunk_beams) *a.error_flag = 1;
  /* Tiles this CTA's 128 consecutive beams touch (they keep hitting the same ones) and which of them hold a beam's own
   * start cell.  The global first-touch probe and the work-list append are done once per CTA and tile AFTER the beams
   * are binned, one lane per tile (HimmArgs::defer_first_touch): inside the loop each probe is a dependent global
   * round trip that stalls its whole warp (10 % of the kernel's stall samples at C4: 0.052 -> 0.047 ms).  A launch
   * that fills the GPU many times over hides that latency anyway and probes at once (C5: 0.430 against 0.455 ms). */
  __shared__ uint32_t s_probed[HIMM_PREP_PROBE_WORDS], s_heavy[HIMM_PREP_PROBE_WORDS];
  const int probe_words = min(HIMM_PREP_PROBE_WORDS, (a.tiles_r * a.tiles_c + 31) >> 5);
  const bool fits = a.tiles_r * a.tiles_c <= 32 * HIMM_PREP_PROBE_WORDS; /* the bitmap covers the robot's tiles */
  const bool deferred = fits && a.defer_first_touch != 0;
  for (int w = threadIdx.x; w < probe_words; w += blockDim.x) {
    s_probed[w] = 0u;
    s_heavy[w] = 0u;
  }
  __syncthreads();
  /* a warp wholly beyond the robot's beams bins nothing (it still joins the barrier below) */
  const bool warp_live = beg + (int)(blockIdx.x * blockDim.x) + (threadIdx.x & ~31) < end;
  const bool valid = warp_live && i < end;
  BeamSeg b;
  b.r0 = b.c0 = b.r1 = b.c1 = b.mr = b.mc = -1;
  b.S = b.B = 0u;
  if (valid) {
    const RobotGeom g = a.geom[a.robot0 + rel];
    double sx, sy, ex, ey;
    int clear_end;
    if (a.samples) {
      const b200nav_sample s = a.samples[i];
      sx = s.sx;
      sy = s.sy;
      ex = s.ex;
      ey = s.ey;
      clear_end = s.clear_end;
    } else if (a.scan_ranges) { /* scan form: project the reading (scan_project.h); dropped readings do nothing */
      const double* pose = a.scan_poses + 3 * (size_t)rel;
      sx = pose[0];
      sy = pose[1];
      double syaw, cyaw;
      b200nav_sincos(pose[2], syaw, cyaw);
      const int j = i - beg;
      const int src = a.scan_sel ? __ldg(&a.scan_sel[j]) : j;
      const float r = __ldg(&a.scan_ranges[(size_t)rel * a.scan.n_ranges + src]);
      clear_end = (a.scan.decimated && scan_clear_end(a.scan, __ldg(&a.scan_ranges[(size_t)rel * a.scan.n_ranges + j]))) ? 1 : 0;
      if (!project_reading(a.scan, j, r, sx, sy, cyaw, syaw, ex, ey)) ex = ey = __longlong_as_double(0x7ff8000000000000ll);
    } else { /* cloud form: Position(*itX, *itY) widens the float32 cloud point (laser_map_updater.cpp:60) */
      const float2 p = a.xy[i];
      sx = a.origins[2 * rel];
      sy = a.origins[2 * rel + 1];
      ex = f32_to_f64(p.x);
      ey = f32_to_f64(p.y);
      clear_end = a.clear_end ? a.clear_end[i] : 0;
    }
    b = make_beam(a.dims, g, sx, sy, ex, ey, clear_end);
  }

  const int k = valid ? i - beg : 0; /* index of the beam within its robot */
  const int chunk = k / a.chunk_beams;
  bool binning = valid;
  if (valid && chunk >= a.n_chunks) {
    *a.error_flag = 1;
    binning = false;
  }
  const int word = (k - chunk * a.chunk_beams) >> 5;
  const int bitpos = k & 31;
  const int n_tiles = a.tiles_r * a.tiles_c;
  const size_t rc_base = ((size_t)rel * a.n_chunks + chunk) * (size_t)n_tiles;

  /* per-lane iterator over the tiles of my beam: first the mark's tile, then band by band along the driving axis */
  const LineForm f = line_form(b);
  const unsigned den = (unsigned)max(f.den, 1);
  /* the line's fixed-point DDA constants: used below for the band ends (q(t) = high word of B + t S, exact) and
   * handed to the tile kernel in the BeamSeg */
  bool diag = false;
  if (b.r0 >= 0) dda_init((unsigned)f.add, den, b.S, b.B, diag);
  if (valid) a.segs[i] = b;
  const int band1 = (f.m0 + f.sm * f.den) / HIMM_TILE;
  int band = f.m0 / HIMM_TILE;
  bool mark_todo = binning && b.mr >= 0;
  bool line_todo = binning && b.r0 >= 0;
  int nt = 0, nt_hi = -1; /* tiles nt..nt_hi of the current band still to emit */

  while (warp_live) {
    int tr = 0, tc = 0;
    bool have = false;
    if (line_todo) {
      if (nt > nt_hi) { /* enter band `band` */
        const int mlo = band * HIMM_TILE, mhi = mlo + HIMM_TILE - 1;
        int ta = (f.sm > 0) ? (mlo - f.m0) : (f.m0 - mhi);
        int tb = (f.sm > 0) ? (mhi - f.m0) : (f.m0 - mlo);
        ta = max(ta, 0);
        tb = min(tb, f.den);
        const int qa = diag ? ta : (int)(dda_at(b.S, b.B, (unsigned)ta) >> 32); /* == (den/2 + ta * add) / den */
        const int qb = diag ? tb : (int)(dda_at(b.S, b.B, (unsigned)tb) >> 32);
        const int na = f.n0 + f.sn * qa, nb = f.n0 + f.sn * qb; /* minor coordinate at both ends (monotone) */
        nt = min(na, nb) / HIMM_TILE;
        nt_hi = max(na, nb) / HIMM_TILE;
      }
      have = true;
      tr = f.row_major ? band : nt;
      tc = f.row_major ? nt : band;
      nt++;
      if (nt > nt_hi) {
        if (band == band1) line_todo = false;
        else band += f.sm;
      }
    } else if (mark_todo) { /* the mark's tile (usually already covered by the line: one more harmless OR) */
      mark_todo = false;
      have = true;
      tr = b.mr / HIMM_TILE;
      tc = b.mc / HIMM_TILE;
    }
    if (__ballot_sync(0xffffffffu, have) == 0u) break;
    const int tile_id = tc * a.tiles_r + tr;
    /* All 32 beams of a warp share the mask word (a warp is 32 consecutive beams, bit = lane), so the lanes that
     * emit the same tile in this iteration own exactly the bits of one word: match.any groups them and the lowest
     * lane of each group issues ONE reduction with the group's lane mask.  Idle lanes share one key (match.any
     * costs issue time per distinct key). */
    const unsigned group = __match_any_sync(0xffffffffu, have ? tile_id : -1);
    const bool head = have && (group & ((1u << lane) - 1u)) == 0u;
    if (head) {
      const size_t widx = (rc_base + (size_t)tile_id) * a.mask_words + word;
      const uint32_t bits = group;
      atomicOr(&a.beam_masks[widx], bits);
      /* The tile that holds the beams' own start cell sees every beam of the scan: such heavy items are queued
       * from the front of the work list, all others from the back, so the long items start first (no long tail). */
      const bool heavy = b.r0 >= 0 && tr == b.r0 / HIMM_TILE && tc == b.c0 / HIMM_TILE;
      if (deferred) {
        const uint32_t bit = 1u << (tile_id & 31);
        atomicOr(&s_probed[tile_id >> 5], bit);
        if (heavy) atomicOr(&s_heavy[tile_id >> 5], bit);
      } else { /* probe at once, but only the first group of this CTA that meets the tile */
        bool probe = true;
        if (fits) {
          const uint32_t bit = 1u << (tile_id & 31);
          probe = (atomicOr(&s_probed[tile_id >> 5], bit) & bit) == 0u;
        }
        if (probe) himm_first_touch(a, rel * n_tiles + tile_id, heavy);
      }
    }
  }
  if (!deferred) return;
  __syncthreads();
  /* first touch of a (robot, tile) in this update appends it to the work list: 32 bitmap words per warp and step, one
   * lane per tile of every non-empty word */
  const int warp = threadIdx.x >> 5;
  for (int base = 32 * warp; base < probe_words; base += (int)blockDim.x) {
    const uint32_t mine = (base + lane < probe_words) ? s_probed[base + lane] : 0u;
    for (unsigned nz = __ballot_sync(0xffffffffu, mine != 0u); nz; nz &= nz - 1u) {
      const int src = __ffs(nz) - 1;
      const uint32_t word = __shfl_sync(0xffffffffu, mine, src);
      if ((word >> lane) & 1u)
        himm_first_touch(a, rel * n_tiles + 32 * (base + src) + lane, ((s_heavy[base + src] >> lane) & 1u) != 0u);
    }
  }
}

/* Work statistics over the BeamSegs of one update: cell visits, marks, beams (block reduce + 3 atomics per CTA). */
__global__ void __launch_bounds__(256) himm_stats_kernel(const BeamSeg* __restrict__ segs, int total,
                                                         unsigned long long* __restrict__ out3) {
  unsigned long long visits = 0, marks = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const BeamSeg b = segs[i];
    if (b.r0 >= 0) visits += (unsigned long long)(max(abs(b.r1 - b.r0), abs(b.c1 - b.c0)) + 1);
    if (b.mr >= 0) marks += 1;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    visits += __shfl_xor_sync(0xffffffffu, visits, o);
    marks += __shfl_xor_sync(0xffffffffu, marks, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&out3[0], visits);
    atomicAdd(&out3[1], marks);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&out3[2], (unsigned long long)total);
}

/* clearCell (map_updater.h:61-71).  v - 10.0 is evaluated in double by the reference; for every float v the
 * float subtraction rounds identically (the double difference is exact for |v| < 2^31 and rounds back to v above). */
__device__ __forceinline__ float himm_clear(float v) {
  v = (v > 0.0f) ? (v - 10.0f) : 0.0f; /* NaN and v <= 0 -> 0 */
  return (v < 0.0f) ? 0.0f : v;
}

/* markCell (map_updater.h:52-59). */
__device__ __forceinline__ float himm_mark(float v) {
  if (!(v > 0.0f)) return 30.0f;
  return (v <= 150.0f) ? (v + 30.0f) : v;
}

/* ---------------------------------------------------------------------------------------------------------------
 * K1: tile-owner update kernel.  One WARP (= one CTA of 32 threads) owns one SUB x SUB tile of one robot's grid.
 *
 * Cell storage in shared memory: HIMM values live in the closed set {NaN, 0, 10, ..., 180} (clearCell / markCell map
 * the set into itself), so a staged tile holds one BYTE per cell: code 0 = NaN, code k = value/10 + 1 (cells.cuh).  That is 4x less
 * shared memory than floats => ~26 resident warps per SM instead of 10, which is what hides the dependent
 * LDS -> op -> STS latency of the in-order walk.  Tiles are converted on load / store (float layers in HBM keep the
 * grid_map layout).  A tile that holds any other value (foreign data uploaded by the host) is processed by the same
 * code through a float view directly on global memory: slower, still exact.
 *   SUB        tile edge (cells); byte pitch SUB+4 -> conflict-free walks along rows and along columns
 *   LIST_CAP   beams per chunk (ordered list of the beams whose bounding box touches the tile)
 * One-warp CTAs let the hardware scheduler balance the very uneven per-tile work (the tile that contains the robot
 * sees every beam) and need no block barriers at all.
 * ------------------------------------------------------------------------------------------------------------- */
template <int SUB, int LIST_CAP>
struct HimmTileCfg {
  static constexpr int kThreads = 32;
  static constexpr int kPitch = SUB + 4; /* bytes; (SUB+4)/4 is odd for SUB = 64 -> column walks hit 32 banks */
  static constexpr int kTileR = SUB;
  static constexpr int kTileC = SUB;
  static constexpr int kTileBytes = SUB * kPitch;
  static constexpr size_t kSmemBytes = kTileBytes + sizeof(uint16_t) * LIST_CAP;
  static_assert(SUB == 64, "tile edge is 64 (two rows per lane, 64-bit column masks)");
};

/* Tile views: the walk is written once against this interface.
 *   peek(off)        the cell as stored; >= 2 means the result of visiting it depends on how many beams visit it and
 *                    in which order (code <= 1, i.e. NaN or 0: any number of clears gives 0)
 *   set_free(off)    result of >= 1 clears on a cell that is not sensitive in that sense
 *   clear_n_known / clear_seq_known   the same on a cell whose content the caller read in this ring block
 *   clear_n / clear_seq   exact application of a group of visits */
/* Shared-memory byte access through an explicit 32-bit shared address (a generic pointer would cost an address
 * space conversion on every access of the hot loop). */
__device__ __forceinline__ int lds_u8(uint32_t addr) {
  int v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_u8(uint32_t addr, int v) {
  asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

struct CodeView { /* shared memory, one byte per cell; offsets are shared-space byte addresses (bias()) */
  uint32_t base; /* shared-space address of the tile */
  __device__ __forceinline__ int bias() const { return (int)base; }
  __device__ __forceinline__ int peek(int off) const { return lds_u8(off); }
  __device__ __forceinline__ void set_free(int off) const { sts_u8(off, 1); }
  /* clear_n / clear_seq on a cell whose content `c` the caller has already read (and nobody wrote since) */
  __device__ __forceinline__ void clear_n_known(int off, int c, int n) const { sts_u8(off, max(c - n, 1)); }
  __device__ __forceinline__ void clear_seq_known(int off, int c, unsigned group, unsigned marks) const {
    while (group) {
      const unsigned bit = group & (0u - group);
      group ^= bit;
      c = max(c - 1, 1);
      if (marks & bit) c = (c <= 16) ? c + 3 : c;
    }
    sts_u8(off, c);
  }
  __device__ __forceinline__ void clear_n(int off, int n, bool mark) const {
    int c = max(lds_u8(off) - n, 1);
    if (mark) c = (c <= 16) ? c + 3 : c;
    sts_u8(off, c);
  }
  /* clears and marks of several beams on one cell, in lane order: lanes in `group`, those in `marks` also mark */
  __device__ __forceinline__ void clear_seq(int off, unsigned group, unsigned marks) const {
    int c = lds_u8(off);
    while (group) {
      const unsigned bit = group & (0u - group);
      group ^= bit;
      c = max(c - 1, 1);
      if (marks & bit) c = (c <= 16) ? c + 3 : c;
    }
    sts_u8(off, c);
  }
  __device__ __forceinline__ void mark(int off) const {
    const int c = lds_u8(off);
    sts_u8(off, (c <= 1) ? 4 : ((c <= 16) ? c + 3 : c));
  }
};
struct FloatView { /* global memory, in place (tiles with values outside the HIMM set) */
  volatile float* p;
  __device__ __forceinline__ int bias() const { return 0; }
  __device__ __forceinline__ int peek(int) const { return 2; }
  __device__ __forceinline__ void set_free(int off) const { p[off] = 0.0f; }
  __device__ __forceinline__ void clear_n_known(int off, int, int n) const { clear_n(off, n, false); }
  __device__ __forceinline__ void clear_seq_known(int off, int, unsigned group, unsigned marks) const {
    clear_seq(off, group, marks);
  }
  __device__ __forceinline__ void clear_n(int off, int n, bool mark) const {
    float v = p[off];
    for (int i = 0; i < n; i++) v = himm_clear(v);
    if (mark) v = himm_mark(v);
    p[off] = v;
  }
  __device__ __forceinline__ void clear_seq(int off, unsigned group, unsigned marks) const {
    float v = p[off];
    while (group) {
      const unsigned bit = group & (0u - group);
      group ^= bit;
      v = himm_clear(v);
      if (marks & bit) v = himm_mark(v);
    }
    p[off] = v;
  }
  __device__ __forceinline__ void mark(int off) const { p[off] = himm_mark(p[off]); }
};

/* Apply the listed beams, in order, to one tile through `view` (cell (r,c) of the tile lives at
 * view[(c-C0)*pitch + (r-R0)]).  32 list entries per batch; see the schedule description below. */
/* What is known about a tile before its beams are applied: which of its 8 x 8 blocks of 8 x 8 cells hold nothing but
 * free cells (value 0; bit bc * 8 + br, block column bc, block row br; the per-tile summary the tile kernel keeps in
 * HimmArgs::free_cols), minus the blocks a mark of an earlier beam of this update has landed in.  A beam segment whose
 * end cells' block box lies in such blocks, and that does not mark in this tile, only re-clears free cells (a
 * Bresenham line is monotone, so all its cells lie in that box): it is dropped before the walk.  Exact: clearing a
 * free cell leaves it free, and only a mark can make a free cell non-free during an update. */
struct FreeBlocks {
  uint32_t lo, hi;             /* summary as loaded for this work item (block columns 0..3 / 4..7) */
  uint32_t dirty_lo, dirty_hi; /* blocks marked by the batches applied so far (this one included) */
  int batches, dropped;        /* statistics: 32-beam batches seen / dropped entirely */
};

template <class View>
__device__ __forceinline__ void himm_apply_list(const View view, const int pitch, const BeamSeg* __restrict__ segs,
                                                const uint16_t* list, const int n_list, const int R0, const int R1,
                                                const int C0, const int C1, const int lane,
                                                FreeBlocks* fb = nullptr) {
  /* Lane-parallel set-up: lane L clips beam L of the batch to this tile and derives the Bresenham state at its
   * first step inside.  Then one of two exact schedules:
   *  (fan)     all beams of the batch start in the SAME cell (a lidar scan).  A cell at step t of such a line has
   *            Chebyshev distance exactly t from that cell, so two beams can only share a cell at EQUAL step index.
   *            Lane L walks its own beam and all lanes advance over t together, so all visits of a cell by this
   *            batch fall into one iteration and are resolved there, in lane order == sample order (see below).
   *  (general) anything else (clipped rays, mixed origins, mark without a line): one beam at a time, the 32 lanes
   *            striding over its cells (a Bresenham line never visits a cell twice). */
  BeamSeg nb;
  nb.r0 = -1;
  nb.mr = -1;
  if (lane < n_list) nb = segs[list[lane]];
  for (int j0 = 0; j0 < n_list; j0 += 32) {
    const int j = j0 + lane;
    const BeamSeg b = nb;
    const bool have = j < n_list;
    /* prefetch the next batch's segments (hides the L2 latency behind this batch's walk) */
    nb.r0 = -1;
    nb.mr = -1;
    if (j + 32 < n_list) nb = segs[list[j + 32]];
    int my_len = 0, my_t0 = 0, my_off0 = view.bias(), my_rem0 = 0, my_dm = 0, my_dn = 0, my_add = 0, my_den = 1, my_moff = -1;
    int my_r0 = -1, my_c0 = -1, my_q0 = 0;
    unsigned my_S = 0u, my_B = 0u; /* fixed-point slope / half offset (dda_init) */
    bool my_diag = false;
    bool mark_at_end = false; /* the mark cell is the last cell of my segment */
    const bool has_mark = have && b.mr >= R0 && b.mr <= R1 && b.mc >= C0 && b.mc <= C1;
    uint32_t eff_lo = 0u, eff_hi = 0u; /* blocks that are free and stay free while this batch is applied */
    if (fb) { /* warp-uniform */
      uint32_t mk_lo = 0u, mk_hi = 0u;
      if (has_mark) {
        const int bit = ((b.mc - C0) >> 3) * 8 + ((b.mr - R0) >> 3);
        if (bit < 32) mk_lo = 1u << bit;
        else mk_hi = 1u << (bit - 32);
      }
      fb->dirty_lo |= __reduce_or_sync(0xffffffffu, mk_lo);
      fb->dirty_hi |= __reduce_or_sync(0xffffffffu, mk_hi);
      eff_lo = fb->lo & ~fb->dirty_lo;
      eff_hi = fb->hi & ~fb->dirty_hi;
    }
    if (have) {
      if (has_mark) my_moff = view.bias() + (b.mc - C0) * pitch + (b.mr - R0);
      if (b.r0 >= 0) {
        const LineForm f = line_form(b);
        int t0, t1;
        if (clip_line_to_rect(f, R0, R1, C0, C1, t0, t1)) {
          const unsigned den = (unsigned)max(f.den, 1);
          my_S = b.S; /* dda_init(add, den), computed once per beam by the binning kernel */
          my_B = b.B;
          my_diag = (unsigned)f.add >= den;
          const unsigned x0 = (unsigned)(f.den >> 1) + (unsigned)t0 * (unsigned)f.add;
          const unsigned q0 = my_diag ? (unsigned)t0 : (unsigned)(dda_at(my_S, my_B, (unsigned)t0) >> 32); /* == x0 / den */
          my_q0 = my_diag ? 0 : (int)q0;
          my_rem0 = (int)(x0 - q0 * den);
          const int mj = f.m0 + f.sm * t0, mn = f.n0 + f.sn * (int)q0;
          const int r = f.row_major ? mj : mn, c = f.row_major ? mn : mj;
          my_off0 = view.bias() + (c - C0) * pitch + (r - R0);
          my_dm = f.row_major ? f.sm : f.sm * pitch;
          my_dn = f.row_major ? f.sn * pitch : f.sn;
          my_add = f.add;
          my_den = (int)den;
          my_len = t1 - t0 + 1;
          my_t0 = t0;
          my_r0 = b.r0;
          my_c0 = b.c0;
          mark_at_end = has_mark && t1 == f.den && b.r1 == b.mr && b.c1 == b.mc;
          if (!has_mark && (eff_lo | eff_hi) != 0u) {
            /* block box of the segment's two end cells against the known-free blocks (see FreeBlocks) */
            const unsigned q1 = my_diag ? (unsigned)t1 : (unsigned)(dda_at(my_S, my_B, (unsigned)t1) >> 32);
            const int mj1 = f.m0 + f.sm * t1, mn1 = f.n0 + f.sn * (int)q1;
            const int re = f.row_major ? mj1 : mn1, ce = f.row_major ? mn1 : mj1;
            const int bra = (min(r, re) - R0) >> 3, brb = (max(r, re) - R0) >> 3;
            const int bca = (min(c, ce) - C0) >> 3, bcb = (max(c, ce) - C0) >> 3;
            const uint32_t rep = (((2u << (brb - bra)) - 1u) << bra) * 0x01010101u; /* the block rows, in every byte */
            uint32_t need_lo = 0u, need_hi = 0u;
            if (bca <= 3) need_lo = rep & (0xffffffffu >> (8 * (3 - min(bcb, 3)))) & (0xffffffffu << (8 * bca));
            if (bcb >= 4) need_hi = rep & (0xffffffffu >> (8 * (7 - bcb))) & (0xffffffffu << (8 * (max(bca, 4) - 4)));
            if (((need_lo & ~eff_lo) | (need_hi & ~eff_hi)) == 0u) my_len = 0; /* only re-clears free cells */
          }
        }
      }
    }
    const bool has_work = my_len > 0 || my_moff >= 0;
    unsigned active = __ballot_sync(0xffffffffu, has_work);
    if (fb) {
      fb->batches++;
      if (active == 0u) fb->dropped++;
    }
    if (active == 0u) continue;
    /* fan test: every lane with work has a segment, marks sit on segment ends, one common start cell */
    const int lead = __ffs(active) - 1;
    const int lr0 = __shfl_sync(0xffffffffu, my_r0, lead), lc0 = __shfl_sync(0xffffffffu, my_c0, lead);
    const bool lane_ok = !has_work || (my_len > 0 && my_r0 == lr0 && my_c0 == lc0 && (my_moff < 0 || mark_at_end));
    if (__all_sync(0xffffffffu, lane_ok)) {
      /* ---- fan schedule: all lanes advance together over the absolute step index t, so every visit this
       * batch pays to a cell happens in ONE iteration.  Cells holding NaN or 0 end up 0 however many beams clear
       * them, so an iteration in which no lane sees a larger value and no lane marks needs no coordination at
       * all (the common case: free space).  Otherwise match.any groups the lanes by cell and the lowest lane of
       * each group applies the group's clears and +30 marks in lane order == sample order. ---- */
      const int first = (my_len > 0) ? my_t0 : 0x7fffffff;
      const unsigned span = (my_len > 0) ? (unsigned)(my_len - 1) : 0u;
      const int last = (my_len > 0) ? my_t0 + my_len - 1 : -0x7fffffff;
      const int tmin = __reduce_min_sync(0xffffffffu, first), tmax = __reduce_max_sync(0xffffffffu, last);
      const int step_plain = my_diag ? my_dm + my_dn : my_dm, step_carry = step_plain + my_dn;
      const int mark_k = (my_moff >= 0) ? (int)span : -1; /* step (relative to `first`) that also marks */
      /* RINGS steps (disjoint rings of cells) per block.  Cells of different steps are different cells, so only the
       * reads and writes of ONE block can meet: a single warp barrier between its read half and its write half
       * orders them.  Stepping is unconditional: frac += S, the carry is the minor step; cells before a line enters
       * the tile or after it left are never dereferenced. */
      constexpr int RINGS = 4;
      auto ring_block = [&](const int k, int& off, unsigned& frac) {
        int offs[RINGS];
        /* my cell of ring r as read: a value >= 2 counts visits, 0 (NaN) becomes free on any visit, 1 is free already
         * and needs no store; -1 when I am not on the ring */
        int val[RINGS];
        bool sens = mark_k >= 0 && (unsigned)(mark_k - k) < (unsigned)RINGS; /* I mark inside this block */
#pragma unroll
        for (int r = 0; r < RINGS; r++) {
          offs[r] = off;
          const bool on = (unsigned)(k + r) <= span;
          val[r] = on ? view.peek(off) : -1;
          sens = sens || val[r] >= 2;
          const unsigned nf = frac + my_S;
          off += (nf < frac) ? step_carry : step_plain;
          frac = nf;
        }
        __syncwarp(); /* all reads of this block are done before any lane writes (memory-model order) */
        if (!__any_sync(0xffffffffu, sens)) {
#pragma unroll
          for (int r = 0; r < RINGS; r++)
            if (val[r] == 0) view.set_free(offs[r]);
        } else {
          /* match.any costs issue time in proportion to the number of DISTINCT keys in the warp (measured on B200:
           * ~4 cycles for one key, 31-83 for 32): lanes that are not on the ring share one key. */
#pragma unroll
          for (int r = 0; r < RINGS; r++) {
            const bool on = val[r] >= 0, marking = k + r == mark_k;
            if (__any_sync(0xffffffffu, val[r] >= 2 || (on && marking))) {
              const unsigned group = __match_any_sync(0xffffffffu, on ? offs[r] : -1);
              const unsigned marks = __ballot_sync(0xffffffffu, on && marking) & group;
              if (on && (group & ((1u << lane) - 1u)) == 0u) {
                if (marks == 0u) view.clear_n_known(offs[r], val[r], __popc(group));
                else view.clear_seq_known(offs[r], val[r], group, marks);
              }
            } else if (val[r] == 0) {
              view.set_free(offs[r]);
            }
          }
        }
      };
      {
        /* every lane starts at step tmin of ITS line and walks on from there */
        const int back = (my_len > 0) ? my_t0 - tmin : 0;
        const unsigned long long xs = dda_at(my_S, my_B, (unsigned)((my_len > 0) ? tmin : 0));
        unsigned frac = (unsigned)xs;
        const int qs = my_diag ? 0 : (int)(xs >> 32);
        int off = my_off0 - back * step_plain - (my_q0 - qs) * my_dn;
        int k = (my_len > 0) ? tmin - first : -0x40000000;
        const int n_iter = (tmax - tmin) / RINGS;
        for (int i = 0; i <= n_iter; i++, k += RINGS) ring_block(k, off, frac);
      }
      __syncwarp(); /* the next batch may read any cell this one wrote */
    } else {
      /* ---- general schedule ---- */
      while (active) {
        const int src = __ffs(active) - 1;
        active &= active - 1;
        const int len = __shfl_sync(0xffffffffu, my_len, src);
        const int moff = __shfl_sync(0xffffffffu, my_moff, src);
        if (len > 0) {
          const int off0 = __shfl_sync(0xffffffffu, my_off0, src);
          const int rem0 = __shfl_sync(0xffffffffu, my_rem0, src);
          const int dm = __shfl_sync(0xffffffffu, my_dm, src);
          const int dn = __shfl_sync(0xffffffffu, my_dn, src);
          const int add = __shfl_sync(0xffffffffu, my_add, src);
          const int den = __shfl_sync(0xffffffffu, my_den, src);
          /* lane L starts at step t0+L: floor((rem0 + L*add)/den) <= 32, exact through a float reciprocal */
          const float rcp = __frcp_rn((float)den);
          const int x = rem0 + lane * add;
          const int q = small_quotient(x, rcp);
          int rem = x - q * den;
          int off = off0 + lane * dm + q * dn;
          const int x32 = 32 * add;
          const int q32 = small_quotient(x32, rcp);
          const int r32 = x32 - q32 * den;
          const int step = 32 * dm + q32 * dn;
          for (int k = lane; k < len; k += 32) {
            view.clear_n(off, 1, false);
            rem += r32;
            off += step;
            if (rem >= den) {
              rem -= den;
              off += dn;
            }
          }
          __syncwarp();
        }
        if (moff >= 0) {
          if (lane == 0) view.mark(moff);
          __syncwarp();
        }
      }
    }
  }
}

template <int SUB, int LIST_CAP>
__global__ void __launch_bounds__(32, 30) himm_tile_kernel(HimmArgs a) {
  using Cfg = HimmTileCfg<SUB, LIST_CAP>;
  static_assert(SUB == HIMM_TILE && LIST_CAP == HIMM_CHUNK, "tile / chunk constants");
  extern __shared__ __align__(128) unsigned char himm_smem_raw[];
  uint8_t* tile = himm_smem_raw;
  uint16_t* list = reinterpret_cast<uint16_t*>(himm_smem_raw + Cfg::kTileBytes); /* a.chunk_beams entries */

  /* shared-space address of the tile, pinned in a register (keeps the address computation out of the hot loop) */
  uint32_t tile_saddr = (uint32_t)__cvta_generic_to_shared(tile);
  asm volatile("mov.u32 %0, %0;" : "+r"(tile_saddr));

  const int lane = threadIdx.x;
  const int rows = a.dims.rows, cols = a.dims.cols;
  const int n_tiles = a.tiles_r * a.tiles_c;
  /* float4 path: every column segment of a tile is 16-byte aligned and a whole number of quads */
  const bool vec_ok = (rows & 3) == 0 && ((reinterpret_cast<uintptr_t>(a.layer) & 15) == 0);
  const int n_heavy = *reinterpret_cast<volatile int*>(&a.counters[0]); /* final: the prep kernel has completed */
  const int n_work = n_heavy + *reinterpret_cast<volatile int*>(&a.counters[3]);

  /* ---- persistent loop: warps pull (robot, tile) work items until the list is empty ---- */
  for (;;) {
  int w = 0;
  if (lane == 0) w = atomicAdd(&a.counters[1], 1);
  w = __shfl_sync(0xffffffffu, w, 0);
  if (w >= n_work) break;
  const int rt = a.worklist[w < n_heavy ? w : a.worklist_cap - 1 - (w - n_heavy)];
  const int rel = rt / n_tiles, tile_id = rt - rel * n_tiles;
  const int robot = a.robot0 + rel;
  const int tile_r = tile_id % a.tiles_r, tile_c = tile_id / a.tiles_r;

  /* tile rectangle (inclusive, clipped to the grid) */
  const int R0 = tile_r * SUB, C0 = tile_c * SUB;
  const int R1 = min(R0 + SUB, rows) - 1, C1 = min(C0 + SUB, cols) - 1;

  int beg;
  if (a.single_n >= 0) beg = 0;
  else beg = __ldg(&a.offsets[rel]);

  float* gbase = static_cast<float*>(a.layer) + (size_t)robot * rows * cols;
  float* gtile = gbase + (size_t)C0 * rows + R0;
  unsigned long long loaded = 0ull; /* columns staged in shared memory (bit c = column C0+c) */
  bool foreign = false;             /* tile holds values outside the HIMM set -> float view on global memory */
  const bool row_lo_ok = R0 + lane <= R1, row_hi_ok = R0 + lane + 32 <= R1;
  if (lane == 0) a.touched[rt] = 0u;
  const unsigned long long free_known = a.free_cols[rt]; /* warp-uniform load */
  bool can_skip = true;

  for (int chunk = 0; chunk < a.n_chunks; chunk++) {
    const size_t t = ((size_t)rel * a.n_chunks + chunk) * (size_t)n_tiles + tile_id;
    /* ---- this tile's beam set: 2048-bit mask written by the prep kernel; consume and clear it ---- */
    uint32_t* mw = a.beam_masks + t * a.mask_words;
    const uint32_t w0 = (lane < a.mask_words) ? mw[lane] : 0u, w1 = (lane + 32 < a.mask_words) ? mw[lane + 32] : 0u;
    if (__ballot_sync(0xffffffffu, (w0 | w1) != 0u) == 0u) continue;
    if (w0) mw[lane] = 0u;
    if (w1) mw[lane + 32] = 0u;
    /* expand the mask into the ordered beam list: lane L owns words L and L+32 */
    const int p0 = __popc(w0), p1 = __popc(w1);
    int inc0 = p0, inc1 = p1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u0 = __shfl_up_sync(0xffffffffu, inc0, o), u1 = __shfl_up_sync(0xffffffffu, inc1, o);
      if (lane >= o) {
        inc0 += u0;
        inc1 += u1;
      }
    }
    const int tot0 = __shfl_sync(0xffffffffu, inc0, 31);
    const int n_list = tot0 + __shfl_sync(0xffffffffu, inc1, 31);
    {
      int pos = inc0 - p0;
      for (uint32_t w = w0; w; w &= w - 1) list[pos++] = (uint16_t)(32 * lane + __ffs(w) - 1);
      pos = tot0 + inc1 - p1;
      for (uint32_t w = w1; w; w &= w - 1) list[pos++] = (uint16_t)(32 * (lane + 32) + __ffs(w) - 1);
    }
    __syncwarp();
    const BeamSeg* segs = a.segs + beg + chunk * a.chunk_beams;

    /* ---- which columns of the tile may be touched?  Conservative: bounding box of each listed beam intersected
     * with the tile (a superset only costs a few extra column loads / stores of unchanged data). ---- */
    unsigned long long need = 0ull;
    bool any_mark = false;
    for (int j = lane; j < n_list; j += 32) {
      const BeamSeg b = segs[list[j]];
      if (b.r0 >= 0 && max(b.r0, b.r1) >= R0 && min(b.r0, b.r1) <= R1) {
        const int ca = max(min(b.c0, b.c1), C0), cb = min(max(b.c0, b.c1), C1);
        if (ca <= cb) need |= (~0ull >> (63 - (cb - ca))) << (ca - C0);
      }
      if (b.mr >= R0 && b.mr <= R1 && b.mc >= C0 && b.mc <= C1) {
        need |= 1ull << (b.mc - C0);
        any_mark = true;
      }
    }
    {
      const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)need), hi = __reduce_or_sync(0xffffffffu, (unsigned)(need >> 32));
      need = ((unsigned long long)hi << 32) | lo;
    }
    /* ---- free-space shortcut: only clears, only on columns known to hold nothing but 0 -> nothing can change ---- */
    if (can_skip && !__any_sync(0xffffffffu, any_mark) && (need & ~free_known) == 0ull) {
      if (lane == 0) atomicAdd(&a.counters[4], 1); /* statistics: tiles skipped */
      continue;
    }
    if (lane == 0) atomicAdd(&a.counters[5], 1);   /* statistics: tiles processed */
    can_skip = false; /* the tile is being modified from here on: later chunks must not trust free_known */

    if (!foreign) {
      /* ---- stage the newly needed columns: float -> code ---- */
      unsigned long long m = need & ~loaded;
      loaded |= m;
      unsigned bad = 0;
      if (vec_ok) {
        /* 16 lanes x float4 cover one column (64 rows): two columns per warp-wide LDG.128, 4 in flight per lane */
        const int half = lane >> 4, quad = lane & 15;
        const bool quad_ok = R0 + 4 * quad <= R1;
        const uint32_t my_saddr = tile_saddr + 4 * quad;
        const float* my_g = gtile + 4 * quad;
#pragma unroll
        for (int h = 0; h < 2; h++) {
          uint32_t w32 = h ? (uint32_t)(m >> 32) : (uint32_t)m;
          while (w32) {
            int cidx[4];
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
              int ca = -1, cb = -1;
              if (w32) {
                ca = __ffs(w32) - 1 + 32 * h;
                w32 &= w32 - 1;
              }
              if (w32) {
                cb = __ffs(w32) - 1 + 32 * h;
                w32 &= w32 - 1;
              }
              cidx[u] = quad_ok ? (half ? cb : ca) : -1;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
              v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (cidx[u] >= 0) v[u] = *reinterpret_cast<const float4*>(my_g + (unsigned)(cidx[u] * rows));
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
              if (cidx[u] >= 0) {
                const unsigned c0 = himm_encode(v[u].x), c1 = himm_encode(v[u].y), c2 = himm_encode(v[u].z),
                               c3 = himm_encode(v[u].w);
                bad |= ((c0 | c1 | c2 | c3) == 255u);
                const uint32_t packed = c0 | (c1 << 8) | (c2 << 16) | (c3 << 24);
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(my_saddr + cidx[u] * Cfg::kPitch), "r"(packed) : "memory");
              }
            }
          }
        }
        m = 0ull;
      } else {
        while (m) {
          int cidx[8];
          float v0[8], v1[8];
#pragma unroll
          for (int u = 0; u < 8; u++) {
            cidx[u] = m ? __ffsll((long long)m) - 1 : -1;
            if (m) m &= m - 1;
          }
#pragma unroll
          for (int u = 0; u < 8; u++) {
            v0[u] = v1[u] = 0.f;
            if (cidx[u] >= 0) {
              const float* src = gtile + (size_t)cidx[u] * rows + lane;
              if (row_lo_ok) v0[u] = src[0];
              if (row_hi_ok) v1[u] = src[32];
            }
          }
#pragma unroll
          for (int u = 0; u < 8; u++) {
            if (cidx[u] >= 0) {
              const unsigned c0 = himm_encode(v0[u]), c1 = himm_encode(v1[u]);
              bad |= (row_lo_ok && c0 == 255u) || (row_hi_ok && c1 == 255u);
              tile[cidx[u] * Cfg::kPitch + lane] = (uint8_t)c0;
              tile[cidx[u] * Cfg::kPitch + lane + 32] = (uint8_t)c1;
            }
          }
        }
      }
      __syncwarp();
      if (__any_sync(0xffffffffu, bad)) {
        /* Foreign values: nothing has been modified yet in this chunk; columns staged by EARLIER chunks hold
         * HIMM-set values only and are flushed first, then everything continues in place on global memory. */
        foreign = true;
      }
    }
    if (foreign && loaded) {
      unsigned long long m = loaded;
      loaded = 0ull;
      /* columns that failed to encode were staged in this very chunk and are unmodified: skip the flush for any
       * column containing a 255 code (its global data is still the truth). */
      while (m) {
        const int c = __ffsll((long long)m) - 1;
        m &= m - 1;
        const unsigned c0 = tile[c * Cfg::kPitch + lane], c1 = tile[c * Cfg::kPitch + lane + 32];
        const bool col_bad = __any_sync(0xffffffffu, (row_lo_ok && c0 == 255u) || (row_hi_ok && c1 == 255u));
        if (!col_bad) {
          float* dst = gtile + (size_t)c * rows + lane;
          if (row_lo_ok) dst[0] = himm_decode(c0);
          if (row_hi_ok) dst[32] = himm_decode(c1);
        }
      }
      __threadfence_block();
      __syncwarp();
    }

    /* ---- apply the beams in sample order ---- */
    if (!foreign)
      himm_apply_list(CodeView{tile_saddr}, Cfg::kPitch, segs, list, n_list, R0, R1, C0, C1, lane);
    else
      himm_apply_list(FloatView{gtile}, rows, segs, list, n_list, R0, R1, C0, C1, lane);
    __syncwarp(); /* the list is rewritten by the next chunk */
  }

  /* ---- write back the staged (== possibly touched) columns: code -> float, coalesced 256-byte segments; and
   * refresh the free-column summary of exactly those columns ---- */
  if (loaded || foreign) {
    unsigned long long m = loaded, now_free = 0ull;
    if (vec_ok) {
      const int half = lane >> 4, quad = lane & 15;
      const bool quad_ok = R0 + 4 * quad <= R1;
      const uint32_t my_saddr = tile_saddr + 4 * quad;
      float* my_g = gtile + 4 * quad;
#pragma unroll
      for (int h = 0; h < 2; h++) {
        uint32_t w32 = h ? (uint32_t)(m >> 32) : (uint32_t)m;
        while (w32) {
          int ca = __ffs(w32) - 1, cb = -1;
          w32 &= w32 - 1;
          if (w32) {
            cb = __ffs(w32) - 1;
            w32 &= w32 - 1;
          }
          const int c = (half ? cb : ca);
          bool quad_free = true; /* rows outside the grid count as free */
          if (c >= 0 && quad_ok) {
            const int cc = c + 32 * h;
            uint32_t w;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(my_saddr + cc * Cfg::kPitch) : "memory");
            quad_free = (w == 0x01010101u);
            float4 o;
            o.x = himm_decode(w & 0xffu);
            o.y = himm_decode((w >> 8) & 0xffu);
            o.z = himm_decode((w >> 16) & 0xffu);
            o.w = himm_decode(w >> 24);
            *reinterpret_cast<float4*>(my_g + (unsigned)(cc * rows)) = o;
          }
          const unsigned fb = __ballot_sync(0xffffffffu, quad_free);
          if ((fb & 0xffffu) == 0xffffu) now_free |= 1ull << (ca + 32 * h);
          if (cb >= 0 && (fb >> 16) == 0xffffu) now_free |= 1ull << (cb + 32 * h);
        }
      }
    } else {
      while (m) {
        const int c = __ffsll((long long)m) - 1;
        m &= m - 1;
        float* dst = gtile + (size_t)c * rows + lane;
        const unsigned c0 = tile[c * Cfg::kPitch + lane], c1 = tile[c * Cfg::kPitch + lane + 32];
        if (row_lo_ok) dst[0] = himm_decode(c0);
        if (row_hi_ok) dst[32] = himm_decode(c1);
        if (__all_sync(0xffffffffu, (!row_lo_ok || c0 == 1u) && (!row_hi_ok || c1 == 1u))) now_free |= 1ull << c;
      }
    }
    /* columns not staged keep what was known about them; a tile processed in place (foreign values) knows nothing */
    const unsigned long long upd = foreign ? 0ull : ((free_known & ~loaded) | now_free);
    if (lane == 0 && upd != free_known) a.free_cols[rt] = upd;
  }
  __syncwarp(); /* the tile buffer is reused by the next work item */
  } /* persistent loop */

  /* the last warp to finish re-arms the counters for the next update */
  if (lane == 0) {
    __threadfence();
    if (atomicAdd(&a.counters[2], 1) == (int)gridDim.x - 1) {
      a.counters[0] = 0;
      a.counters[1] = 0;
      a.counters[2] = 0;
      a.counters[3] = 0;
    }
  }
}

/* ---------------------------------------------------------------------------------------------------------------
 * K1 for CODED layers (cells.cuh): the tile record in HBM is the shared-memory image, so a work item is
 *   one bulk async copy in (cp.async.bulk global -> shared, completion on an mbarrier), the in-order walk, one bulk
 *   async copy out (shared -> global).  No conversion, no per-column address arithmetic, 4352 bytes each way.
 * Free-tile shortcut: a record whose cells are all 0 (code 1) is remembered in free_cols[] (all ones); while that
 * holds, a beam set without marks cannot change the tile and the work item is dropped before any copy.
 * ------------------------------------------------------------------------------------------------------------- */
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(mbar), "r"(parity)
        : "memory");
  }
}

/* Resident CTAs per SM the register allocation is sized for: 28 (the launch's default grid, b200nav_api.cu) leaves 72
 * registers per thread and no spills; 30 caps at 64 with a few spills (measured 0.164 -> 0.161 ms at C4). */
#ifndef B200NAV_TILE_MIN_BLOCKS
#define B200NAV_TILE_MIN_BLOCKS 28
#endif
template <int LIST_CAP>
__global__ void __launch_bounds__(32, B200NAV_TILE_MIN_BLOCKS) himm_tile_coded_kernel(HimmArgs a) {
  static_assert(LIST_CAP == HIMM_CHUNK, "chunk constant");
  extern __shared__ __align__(128) unsigned char himm_smem_raw[];
  uint8_t* tile = himm_smem_raw;
  uint16_t* list = reinterpret_cast<uint16_t*>(himm_smem_raw + HIMM_TILE_BYTES); /* a.chunk_beams entries */
  __shared__ __align__(8) unsigned long long s_mbar;

  uint32_t tile_saddr = (uint32_t)__cvta_generic_to_shared(tile);
  asm volatile("mov.u32 %0, %0;" : "+r"(tile_saddr));
  const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(&s_mbar);

  const int lane = threadIdx.x;
  const int rows = a.dims.rows, cols = a.dims.cols;
  const int n_tiles = a.tiles_r * a.tiles_c;
  const int n_heavy = *reinterpret_cast<volatile int*>(&a.counters[0]); /* final: the prep kernel has completed */
  const int n_light = *reinterpret_cast<volatile int*>(&a.counters[3]);
  const int n_work = n_heavy + n_light;
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t phase = 0; /* mbarrier phase parity of the next copy-in */
  /* Work items are claimed when the warp is free, never ahead: a pre-claimed item waits for its owner while other
   * warps idle, and items differ in length by two orders of magnitude (measured at C4: claiming one ahead 0.153 ->
   * 0.213 ms; one ahead only behind light items 0.163 ms). */
  for (;;) {
    int w = 0;
    if (lane == 0) w = atomicAdd(&a.counters[1], 1);
    w = __shfl_sync(0xffffffffu, w, 0);
    if (w >= n_work) break;
    const int rt = a.worklist[w < n_heavy ? w : a.worklist_cap - 1 - (w - n_heavy)];
    const int rel = rt / n_tiles, tile_id = rt - rel * n_tiles;
    const int robot = a.robot0 + rel;
    const int tile_r = tile_id % a.tiles_r, tile_c = tile_id / a.tiles_r;
    const int R0 = tile_r * HIMM_TILE, C0 = tile_c * HIMM_TILE;
    const int R1 = min(R0 + HIMM_TILE, rows) - 1, C1 = min(C0 + HIMM_TILE, cols) - 1;
    int beg;
    if (a.single_n >= 0) beg = 0;
    else beg = __ldg(&a.offsets[rel]);
    uint8_t* grec = static_cast<uint8_t*>(a.layer) + ((size_t)robot * n_tiles + tile_id) * HIMM_TILE_BYTES;
    if (lane == 0) a.touched[rt] = 0u;
    const unsigned long long fsum = a.free_cols[rt]; /* warp-uniform load: known-free 8 x 8 blocks of the tile */
    const bool known_free = fsum == ~0ull;
    FreeBlocks fb;
    fb.lo = (uint32_t)fsum;
    fb.hi = (uint32_t)(fsum >> 32);
    fb.dirty_lo = fb.dirty_hi = 0u;
    fb.batches = fb.dropped = 0;
    bool staged = false; /* copy-in issued */
    bool ready = false;  /* copy-in complete (waited for) */

    auto stage = [&]() {
      if (lane == 0) {
        /* the previous work item's store must have finished READING the buffer before it is overwritten */
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "n"(HIMM_TILE_BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(tile_saddr), "l"(grec), "n"(HIMM_TILE_BYTES), "r"(mbar)
                     : "memory");
      }
      staged = true;
    };
    if (!known_free) stage(); /* the copy flies while the beam list is being built */

    for (int chunk = 0; chunk < a.n_chunks; chunk++) {
      const size_t t = ((size_t)rel * a.n_chunks + chunk) * (size_t)n_tiles + tile_id;
      /* ---- this tile's beam set: bit mask written by the prep kernel; consume and clear it ---- */
      uint32_t* mw = a.beam_masks + t * a.mask_words;
      const uint32_t w0 = (lane < a.mask_words) ? mw[lane] : 0u, w1 = (lane + 32 < a.mask_words) ? mw[lane + 32] : 0u;
      if (__ballot_sync(0xffffffffu, (w0 | w1) != 0u) == 0u) continue;
      if (w0) mw[lane] = 0u;
      if (w1) mw[lane + 32] = 0u;
      /* expand the mask into the ordered beam list: lane L owns words L and L+32 */
      const int p0 = __popc(w0), p1 = __popc(w1);
      int inc0 = p0, inc1 = p1;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u0 = __shfl_up_sync(0xffffffffu, inc0, o), u1 = __shfl_up_sync(0xffffffffu, inc1, o);
        if (lane >= o) {
          inc0 += u0;
          inc1 += u1;
        }
      }
      const int tot0 = __shfl_sync(0xffffffffu, inc0, 31);
      const int n_list = tot0 + __shfl_sync(0xffffffffu, inc1, 31);
      {
        int pos = inc0 - p0;
        for (uint32_t ww = w0; ww; ww &= ww - 1) list[pos++] = (uint16_t)(32 * lane + __ffs(ww) - 1);
        pos = tot0 + inc1 - p1;
        for (uint32_t ww = w1; ww; ww &= ww - 1) list[pos++] = (uint16_t)(32 * (lane + 32) + __ffs(ww) - 1);
      }
      __syncwarp();
      const BeamSeg* segs = a.segs + beg + chunk * a.chunk_beams;

      if (!staged) {
        /* ---- free-space shortcut: every cell is 0 and no beam of this chunk marks -> nothing can change ---- */
        bool any_mark = false;
        for (int j = lane; j < n_list; j += 32) {
          const BeamSeg b = segs[list[j]];
          any_mark |= b.mr >= R0 && b.mr <= R1 && b.mc >= C0 && b.mc <= C1;
        }
        if (!__any_sync(0xffffffffu, any_mark)) {
          if (lane == 0) atomicAdd(&a.counters[4], 1); /* statistics: tiles skipped */
          continue;
        }
        stage();
      }
      if (lane == 0) atomicAdd(&a.counters[5], 1); /* statistics: tiles processed */
      if (!ready) { /* the first chunk that needs the data waits for the copy */
        mbar_wait(mbar, phase);
        phase ^= 1u;
        ready = true;
      }
      himm_apply_list(CodeView{tile_saddr}, HIMM_TILE_PITCH, segs, list, n_list, R0, R1, C0, C1, lane, &fb);
      __syncwarp(); /* the list is rewritten by the next chunk */
    }

    if (staged) {
      if (!ready) { /* staged but no chunk had beams (stale work item): keep the barrier phase right */
        mbar_wait(mbar, phase);
        phase ^= 1u;
      }
      /* Which 8 x 8 blocks of the record hold nothing but free cells now?  (cells outside the grid hold the free code
       * for good.)  Lane L owns columns 2L and 2L+1: eight block-row bits per column from sixteen words, ANDed over
       * the eight columns (four lanes) of a block column, gathered into the 64-bit summary (bit bc * 8 + br). */
      uint32_t m = 0xffu;
#pragma unroll
      for (int cc = 0; cc < 2; cc++) {
        const uint32_t col = tile_saddr + (uint32_t)((2 * lane + cc) * HIMM_TILE_PITCH);
#pragma unroll
        for (int br = 0; br < 8; br++) {
          uint32_t w0, w1;
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(col + 8 * br) : "memory");
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w1) : "r"(col + 8 * br + 4) : "memory");
          if (((w0 ^ 0x01010101u) | (w1 ^ 0x01010101u)) != 0u) m &= ~(1u << br);
        }
      }
      m &= __shfl_xor_sync(0xffffffffu, m, 1);
      m &= __shfl_xor_sync(0xffffffffu, m, 2);
      const int bc = lane >> 2;
      const uint32_t sum_lo = __reduce_or_sync(0xffffffffu, ((lane & 3) == 0 && bc < 4) ? m << (8 * bc) : 0u);
      const uint32_t sum_hi = __reduce_or_sync(0xffffffffu, ((lane & 3) == 0 && bc >= 4) ? m << (8 * (bc - 4)) : 0u);
      const unsigned long long new_sum = ((unsigned long long)sum_hi << 32) | sum_lo;
      /* generic-proxy writes of the walk -> visible to the async proxy, then one lane sends the record home */
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(grec), "r"(tile_saddr), "n"(HIMM_TILE_BYTES) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (new_sum != fsum) a.free_cols[rt] = new_sum;
      }
    }
    if (lane == 0 && fb.batches) { /* statistics: 32-beam batches seen / dropped because they only re-clear free cells */
      atomicAdd(&a.counters[6], fb.batches);
      if (fb.dropped) atomicAdd(&a.counters[7], fb.dropped);
    }
    __syncwarp();
  } /* persistent loop */

  if (lane == 0) {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); /* all records are home before the CTA retires */
    /* the last warp to finish re-arms the counters for the next update */
    __threadfence();
    if (atomicAdd(&a.counters[2], 1) == (int)gridDim.x - 1) {
      a.counters[0] = 0;
      a.counters[1] = 0;
      a.counters[2] = 0;
      a.counters[3] = 0;
    }
  }
}

/* ---------------------------------------------------------------------------------------------------------------
 * Multi-warp walk of ONE tile's beam list: a wavefront pipeline over the batches.
 *
 * The tile that holds a scan's own origin sees every beam of the scan; walked by one warp it is the critical path of
 * an update.  Here the NW warps of a CTA share the tile: warp w owns batches w, w + NW, w + 2 NW, ... of the ordered
 * beam list and walks each of them exactly as the one-warp kernel does (same set-up, same ring blocks, same exact
 * path) - no work is replicated.  What the reference's order demands is only that, cell by cell, batch j is applied
 * after batch j - 1.  For fan batches that share their origin cell a cell at step t of a line has Chebyshev distance
 * t from the origin, so batch j may enter the block of steps [ts, ts + 4) as soon as batch j - 1 has left it: every
 * warp publishes its progress (batch number, origin, steps completed) in one 64-bit shared-memory word and its
 * successor waits on that word before each block - a systolic wavefront, batch j one block behind batch j - 1.  If
 * the two batches differ in origin, or either of them is not a fan (general schedule), the successor waits until its
 * predecessor has finished.  The set-up of a batch does not touch the tile and runs ahead of the predecessor's walk.
 *   progress word: (batch + 1) << 48 | steps_done << 32 | origin   (steps_done = 0xffff: finished; origin = ~0: no fan)
 * ------------------------------------------------------------------------------------------------------------- */
#define HIMM_PIPE_DONE 0xffffull
#define HIMM_PIPE_NOFAN 0xffffffffull

/* The progress words are the ONLY synchronisation between the warps of a CTA inside a list: a release store after the
 * warp's cell writes (ordered before it by __syncwarp), an acquire load before the successor's cell reads.
 * compute-sanitizer's racecheck does not model synchronisation through memory and reports these words and the cells
 * they guard as hazards (profiles/README.md); the one-warp kernel is racecheck-clean. */
__device__ __forceinline__ unsigned long long pipe_load(uint32_t addr) {
  unsigned long long v;
  asm volatile("ld.acquire.cta.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void pipe_store(uint32_t addr, unsigned long long v) {
  asm volatile("st.release.cta.shared.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}

template <int NW>
__device__ __forceinline__ void himm_apply_list_pipe(const CodeView view, const int pitch,
                                                     const BeamSeg* __restrict__ segs, const uint16_t* list,
                                                     const int n_list, const int R0, const int R1, const int C0,
                                                     const int C1, const int lane, const int warp,
                                                     const uint32_t prog /* shared address of NW progress words */,
                                                     int* err) {
  /* Every wait below is for a warp of the same CTA that makes progress on its own (independent thread scheduling);
   * the bound only exists so that a logic error could never wedge the GPU: it raises the update's error flag. */
  unsigned spins = 0u;
  auto stuck = [&]() {
    if (++spins < (1u << 27)) return false;
    *err = 2;
    return true;
  };
  constexpr int RINGS = 4;
  const uint32_t my_word = prog + 8u * (uint32_t)warp, pred_word = prog + 8u * (uint32_t)((warp + NW - 1) % NW);
  BeamSeg nb;
  nb.r0 = -1;
  nb.mr = -1;
  if (32 * warp + lane < n_list) nb = segs[list[32 * warp + lane]];
  for (int j0 = 32 * warp; j0 < n_list; j0 += 32 * NW) {
    const int jb = j0 >> 5; /* batch number */
    const int j = j0 + lane;
    const BeamSeg b = nb;
    const bool have = j < n_list;
    nb.r0 = -1;
    nb.mr = -1;
    if (j + 32 * NW < n_list) nb = segs[list[j + 32 * NW]];
    /* ---- set-up (as himm_apply_list) ---- */
    int my_len = 0, my_t0 = 0, my_off0 = view.bias(), my_rem0 = 0, my_dm = 0, my_dn = 0, my_add = 0, my_den = 1, my_moff = -1;
    int my_r0 = -1, my_c0 = -1, my_q0 = 0;
    unsigned my_S = 0u, my_B = 0u;
    bool my_diag = false, mark_at_end = false;
    if (have) {
      const bool has_mark = b.mr >= R0 && b.mr <= R1 && b.mc >= C0 && b.mc <= C1;
      if (has_mark) my_moff = view.bias() + (b.mc - C0) * pitch + (b.mr - R0);
      if (b.r0 >= 0) {
        const LineForm f = line_form(b);
        int t0, t1;
        if (clip_line_to_rect(f, R0, R1, C0, C1, t0, t1)) {
          const unsigned den = (unsigned)max(f.den, 1);
          my_S = b.S; /* dda_init(add, den), computed once per beam by the binning kernel */
          my_B = b.B;
          my_diag = (unsigned)f.add >= den;
          const unsigned x0 = (unsigned)(f.den >> 1) + (unsigned)t0 * (unsigned)f.add;
          const unsigned q0 = my_diag ? (unsigned)t0 : (unsigned)(dda_at(my_S, my_B, (unsigned)t0) >> 32);
          my_q0 = my_diag ? 0 : (int)q0;
          my_rem0 = (int)(x0 - q0 * den);
          const int mj = f.m0 + f.sm * t0, mn = f.n0 + f.sn * (int)q0;
          const int r = f.row_major ? mj : mn, c = f.row_major ? mn : mj;
          my_off0 = view.bias() + (c - C0) * pitch + (r - R0);
          my_dm = f.row_major ? f.sm : f.sm * pitch;
          my_dn = f.row_major ? f.sn * pitch : f.sn;
          my_add = f.add;
          my_den = (int)den;
          my_len = t1 - t0 + 1;
          my_t0 = t0;
          my_r0 = b.r0;
          my_c0 = b.c0;
          mark_at_end = has_mark && t1 == f.den && b.r1 == b.mr && b.c1 == b.mc;
        }
      }
    }
    const bool has_work = my_len > 0 || my_moff >= 0;
    unsigned active = __ballot_sync(0xffffffffu, has_work);
    const unsigned long long tag = (unsigned long long)(jb + 1) << 48;
    bool fan = false;
    unsigned long long origin = HIMM_PIPE_NOFAN;
    if (active != 0u) {
      const int lead = __ffs(active) - 1;
      const int lr0 = __shfl_sync(0xffffffffu, my_r0, lead), lc0 = __shfl_sync(0xffffffffu, my_c0, lead);
      const bool lane_ok = !has_work || (my_len > 0 && my_r0 == lr0 && my_c0 == lc0 && (my_moff < 0 || mark_at_end));
      fan = __all_sync(0xffffffffu, lane_ok);
      if (fan) origin = (unsigned long long)(unsigned)(lr0 * 65536 + lc0);
    }
    /* "Finished" is a statement about the whole chain: batch jb may only say it once batch jb - 1 has (a batch that
     * covers few steps, or none, must not let its successor overtake an earlier, longer batch). */
    auto wait_pred_finished = [&]() {
      if (jb == 0) return;
      unsigned long long p;
      do {
        p = pipe_load(pred_word);
      } while (((p >> 48) < (unsigned long long)jb ||
                ((p >> 48) == (unsigned long long)jb && ((p >> 32) & 0xffffull) != HIMM_PIPE_DONE)) &&
               !stuck());
    };
    /* announce the batch (nothing done yet) */
    if (lane == 0) pipe_store(my_word, tag | origin);
    if (active == 0u) { /* an empty batch only passes the chain's state on */
      wait_pred_finished();
      if (lane == 0) pipe_store(my_word, tag | (HIMM_PIPE_DONE << 32) | origin);
      continue;
    }
    /* ---- dependency on batch jb - 1 (owned by the previous warp) ---- */
    bool pred_done = jb == 0;    /* nothing (more) to wait for */
    unsigned pred_steps = 0u;    /* steps [0, pred_steps) of the predecessor are complete (same origin, both fans) */
    if (!pred_done) {
      unsigned long long p;
      do {
        p = pipe_load(pred_word);
      } while ((p >> 48) < (unsigned long long)jb && !stuck()); /* batch jb - 1 not announced yet */
      if ((p >> 48) > (unsigned long long)jb || ((p >> 32) & 0xffffull) == HIMM_PIPE_DONE) {
        pred_done = true;
      } else if (!fan || (p & 0xffffffffull) != origin) {
        do { /* other origin, or one of us is no fan: any cell may be shared at any step */
          p = pipe_load(pred_word);
        } while ((p >> 48) == (unsigned long long)jb && ((p >> 32) & 0xffffull) != HIMM_PIPE_DONE && !stuck());
        pred_done = true;
      } else {
        pred_steps = (unsigned)((p >> 32) & 0xffffull);
      }
      __threadfence_block();
    }
    if (fan) {
      const int first = (my_len > 0) ? my_t0 : 0x7fffffff;
      const unsigned span = (my_len > 0) ? (unsigned)(my_len - 1) : 0u;
      const int last = (my_len > 0) ? my_t0 + my_len - 1 : -0x7fffffff;
      const int tmin = __reduce_min_sync(0xffffffffu, first), tmax = __reduce_max_sync(0xffffffffu, last);
      const int step_plain = my_diag ? my_dm + my_dn : my_dm, step_carry = step_plain + my_dn;
      const int mark_k = (my_moff >= 0) ? (int)span : -1;
      const int back = (my_len > 0) ? my_t0 - tmin : 0;
      const unsigned long long xs = dda_at(my_S, my_B, (unsigned)((my_len > 0) ? tmin : 0));
      unsigned frac = (unsigned)xs;
      const int qs = my_diag ? 0 : (int)(xs >> 32);
      int off = my_off0 - back * step_plain - (my_q0 - qs) * my_dn;
      int k = (my_len > 0) ? tmin - first : -0x40000000;
      const int n_iter = (tmax - tmin) / RINGS;
      for (int i = 0; i <= n_iter; i++, k += RINGS) {
        const unsigned need = (unsigned)(tmin + RINGS * (i + 1)); /* my block covers steps < need */
        if (!pred_done && pred_steps < need) {
          unsigned long long p;
          for (;;) {
            p = pipe_load(pred_word);
            if ((p >> 48) > (unsigned long long)jb || ((p >> 32) & 0xffffull) == HIMM_PIPE_DONE) {
              pred_done = true;
              break;
            }
            pred_steps = (unsigned)((p >> 32) & 0xffffull);
            if (pred_steps >= need || stuck()) break;
          }
          __threadfence_block();
        }
        /* one block of RINGS steps: identical to himm_apply_list's ring_block */
        int offs[RINGS];
        int val[RINGS];
        bool sens = mark_k >= 0 && (unsigned)(mark_k - k) < (unsigned)RINGS;
#pragma unroll
        for (int r = 0; r < RINGS; r++) {
          offs[r] = off;
          const bool on = (unsigned)(k + r) <= span;
          val[r] = on ? view.peek(off) : -1;
          sens = sens || val[r] >= 2;
          const unsigned nf = frac + my_S;
          off += (nf < frac) ? step_carry : step_plain;
          frac = nf;
        }
        __syncwarp();
        if (!__any_sync(0xffffffffu, sens)) {
#pragma unroll
          for (int r = 0; r < RINGS; r++)
            if (val[r] == 0) view.set_free(offs[r]);
        } else {
#pragma unroll
          for (int r = 0; r < RINGS; r++) {
            const bool on = val[r] >= 0, marking = k + r == mark_k;
            if (__any_sync(0xffffffffu, val[r] >= 2 || (on && marking))) {
              const unsigned group = __match_any_sync(0xffffffffu, on ? offs[r] : -1);
              const unsigned marks = __ballot_sync(0xffffffffu, on && marking) & group;
              if (on && (group & ((1u << lane) - 1u)) == 0u) {
                if (marks == 0u) view.clear_n_known(offs[r], val[r], __popc(group));
                else view.clear_seq_known(offs[r], val[r], group, marks);
              }
            } else if (val[r] == 0) {
              view.set_free(offs[r]);
            }
          }
        }
        /* publish: every step below `need` of this batch is complete (the last block publishes "finished" below) */
        if (i < n_iter) {
          __threadfence_block();
          __syncwarp();
          if (lane == 0) pipe_store(my_word, tag | ((unsigned long long)min(need, 0xfffeu) << 32) | origin);
        }
      }
      __syncwarp();
    } else {
      /* ---- general schedule (the predecessor has finished): one beam at a time, lanes striding over its cells ---- */
      while (active) {
        const int src = __ffs(active) - 1;
        active &= active - 1;
        const int len = __shfl_sync(0xffffffffu, my_len, src);
        const int moff = __shfl_sync(0xffffffffu, my_moff, src);
        if (len > 0) {
          const int off0 = __shfl_sync(0xffffffffu, my_off0, src);
          const int rem0 = __shfl_sync(0xffffffffu, my_rem0, src);
          const int dm = __shfl_sync(0xffffffffu, my_dm, src);
          const int dn = __shfl_sync(0xffffffffu, my_dn, src);
          const int add = __shfl_sync(0xffffffffu, my_add, src);
          const int den = __shfl_sync(0xffffffffu, my_den, src);
          const float rcp = __frcp_rn((float)den);
          const int x = rem0 + lane * add;
          const int q = small_quotient(x, rcp);
          int rem = x - q * den;
          int off = off0 + lane * dm + q * dn;
          const int x32 = 32 * add;
          const int q32 = small_quotient(x32, rcp);
          const int r32 = x32 - q32 * den;
          const int step = 32 * dm + q32 * dn;
          for (int kk = lane; kk < len; kk += 32) {
            view.clear_n(off, 1, false);
            rem += r32;
            off += step;
            if (rem >= den) {
              rem -= den;
              off += dn;
            }
          }
          __syncwarp();
        }
        if (moff >= 0) {
          if (lane == 0) view.mark(moff);
          __syncwarp();
        }
      }
    }
    /* finished (and so is everything before me): successors may touch any cell */
    if (!pred_done) wait_pred_finished();
    __threadfence_block();
    __syncwarp();
    if (lane == 0) pipe_store(my_word, tag | (HIMM_PIPE_DONE << 32) | origin);
  }
}

/* ---------------------------------------------------------------------------------------------------------------
 * K1 for small fleets (HimmArgs::mw_all: every touched tile; one warp per tile could not fill the GPU anyway, and the
 * tile that holds a scan's own origin - it sees every beam - would keep a single warp busy for most of the update):
 * NW warps share one tile as a wavefront pipeline over the beam batches (himm_apply_list_pipe).  CTA b handles work
 * items b, b + gridDim.x, ... (the heavy origin tiles first); the last CTA re-arms the counters.
 * ------------------------------------------------------------------------------------------------------------- */
template <int LIST_CAP, int NW>
__global__ void __launch_bounds__(32 * NW) himm_tile_coded_mw_kernel(HimmArgs a) {
  static_assert(LIST_CAP == HIMM_CHUNK, "chunk constant");
  extern __shared__ __align__(128) unsigned char himm_smem_raw[];
  uint8_t* tile = himm_smem_raw;
  uint16_t* list = reinterpret_cast<uint16_t*>(himm_smem_raw + HIMM_TILE_BYTES);
  __shared__ __align__(8) unsigned long long s_mbar;
  __shared__ int s_nlist;
  __shared__ __align__(8) unsigned long long s_prog[NW]; /* pipeline progress words (himm_apply_list_pipe) */

  uint32_t tile_saddr = (uint32_t)__cvta_generic_to_shared(tile);
  asm volatile("mov.u32 %0, %0;" : "+r"(tile_saddr));
  const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(&s_mbar);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rows = a.dims.rows, cols = a.dims.cols;
  const int n_tiles = a.tiles_r * a.tiles_c;
  const int n_heavy = *reinterpret_cast<volatile int*>(&a.counters[0]);
  const int n_items = a.mw_all ? n_heavy + *reinterpret_cast<volatile int*>(&a.counters[3]) : n_heavy;
  const uint32_t prog = (uint32_t)__cvta_generic_to_shared(&s_prog[0]);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t phase = 0;

  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int rt = a.worklist[item < n_heavy ? item : a.worklist_cap - 1 - (item - n_heavy)];
    const int rel = rt / n_tiles, tile_id = rt - rel * n_tiles;
    const int robot = a.robot0 + rel;
    const int tile_r = tile_id % a.tiles_r, tile_c = tile_id / a.tiles_r;
    const int R0 = tile_r * HIMM_TILE, C0 = tile_c * HIMM_TILE;
    const int R1 = min(R0 + HIMM_TILE, rows) - 1, C1 = min(C0 + HIMM_TILE, cols) - 1;
    int beg;
    if (a.single_n >= 0) beg = 0;
    else beg = __ldg(&a.offsets[rel]);
    uint8_t* grec = static_cast<uint8_t*>(a.layer) + ((size_t)robot * n_tiles + tile_id) * HIMM_TILE_BYTES;
    const bool known_free = a.free_cols[rt] == ~0ull;
    if (threadIdx.x == 0) {
      a.touched[rt] = 0u;
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); /* the previous item's store has read the buffer */
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "n"(HIMM_TILE_BYTES) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(tile_saddr), "l"(grec), "n"(HIMM_TILE_BYTES), "r"(mbar)
                   : "memory");
    }
    bool ready = false;

    for (int chunk = 0; chunk < a.n_chunks; chunk++) {
      if (warp == 0) { /* mask -> ordered beam list, as in the one-warp kernel */
        const size_t t = ((size_t)rel * a.n_chunks + chunk) * (size_t)n_tiles + tile_id;
        uint32_t* mw = a.beam_masks + t * a.mask_words;
        const uint32_t w0 = (lane < a.mask_words) ? mw[lane] : 0u, w1 = (lane + 32 < a.mask_words) ? mw[lane + 32] : 0u;
        if (w0) mw[lane] = 0u;
        if (w1) mw[lane + 32] = 0u;
        const int p0 = __popc(w0), p1 = __popc(w1);
        int inc0 = p0, inc1 = p1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int u0 = __shfl_up_sync(0xffffffffu, inc0, o), u1 = __shfl_up_sync(0xffffffffu, inc1, o);
          if (lane >= o) {
            inc0 += u0;
            inc1 += u1;
          }
        }
        const int tot0 = __shfl_sync(0xffffffffu, inc0, 31);
        const int n = tot0 + __shfl_sync(0xffffffffu, inc1, 31);
        int pos = inc0 - p0;
        for (uint32_t ww = w0; ww; ww &= ww - 1) list[pos++] = (uint16_t)(32 * lane + __ffs(ww) - 1);
        pos = tot0 + inc1 - p1;
        for (uint32_t ww = w1; ww; ww &= ww - 1) list[pos++] = (uint16_t)(32 * (lane + 32) + __ffs(ww) - 1);
        if (lane == 0) s_nlist = n;
        if (lane < NW) s_prog[lane] = 0ull; /* no batch of this list announced yet */
      }
      __syncthreads();
      const int n_list = s_nlist;
      if (n_list > 0) {
        if (!ready) {
          mbar_wait(mbar, phase);
          phase ^= 1u;
          ready = true;
        }
        if (threadIdx.x == 0) atomicAdd(&a.counters[5], 1); /* statistics: tiles processed */
        const BeamSeg* segs = a.segs + beg + chunk * a.chunk_beams;
        himm_apply_list_pipe<NW>(CodeView{tile_saddr}, HIMM_TILE_PITCH, segs, list, n_list, R0, R1, C0, C1, lane, warp, prog,
                                 a.error_flag);
      }
      __syncthreads(); /* all warps are done with the list (and with the tile, after the last chunk) */
    }
    if (!ready) {
      mbar_wait(mbar, phase);
      phase ^= 1u;
    }
    unsigned diff = 0;
    for (int i = threadIdx.x; i < HIMM_TILE_BYTES / 16; i += blockDim.x) {
      uint32_t x, y, z, q;
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(q) : "r"(tile_saddr + 16 * i) : "memory");
      diff |= (x ^ 0x01010101u) | (y ^ 0x01010101u) | (z ^ 0x01010101u) | (q ^ 0x01010101u);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const bool all_free = __syncthreads_or(diff != 0u) == 0;
    if (threadIdx.x == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(grec), "r"(tile_saddr), "n"(HIMM_TILE_BYTES) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (all_free != known_free) a.free_cols[rt] = all_free ? ~0ull : 0ull;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (a.mw_all) { /* no one-warp launch follows: the last CTA re-arms the counters for the next update */
      __threadfence();
      if (atomicAdd(&a.counters[2], 1) == (int)gridDim.x - 1) {
        a.counters[0] = 0;
        a.counters[1] = 0;
        a.counters[2] = 0;
        a.counters[3] = 0;
      }
    }
  }
}

}  // namespace b200nav
#endif

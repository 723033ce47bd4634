/*
 * grid_kernels.cuh -- small layer-maintenance kernels (NaN fill, strip fill for GridMap::move, OccupancyGrid export).
 *
 * Replaces (behaviour, not code): GridMap::clearAll / clearRows / clearCols
 * (grid_map_core/src/GridMap.cpp:624-650) and GridMapRosConverter::toOccupancyGrid
 * (grid_map_ros/src/GridMapRosConverter.cpp:251-287).  All HBM-bound, coalesced along rows (column-major layers).
 */
#ifndef B200NAV_GRID_KERNELS_CUH
#define B200NAV_GRID_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "cells.cuh"
#include "geometry.h"

namespace b200nav {

__global__ void grid_fill_kernel(float* __restrict__ p, size_t n, float value) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = value;
}

/* Fill rows [r0, r0+nr) x cols [c0, c0+nc) of one column-major layer. grid = (ceil(nr/128), min(nc, 65535)). */
__global__ void grid_fill_rect_kernel(float* __restrict__ base, int rows, int r0, int nr, int c0, int nc, float value) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nr) return;
  for (int c = blockIdx.y; c < nc; c += gridDim.y) base[(size_t)(c0 + c) * rows + (r0 + r)] = value;
}

/* One thread per buffer cell: value -> int8 [-1, 0..100] written at the reversed unwrapped linear index. */
__global__ void grid_to_occupancy_kernel(const LayerRef layer, int rows, int cols, int start0, int start1,
                                         float data_min, float data_max, int8_t* __restrict__ out) {
  const size_t n = (size_t)rows * cols;
  const size_t lin = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lin >= n) return;
  const int b0 = (int)(lin % rows), b1 = (int)(lin / rows);
  float value = (layer.at(b0, b1) - data_min) / (data_max - data_min);
  if (isnan(value) || (value < 0))
    value = -1;
  else
    value = 0.0f + fminf(fmaxf(0.0f, value), 1.0f) * 100.0f;
  int u0 = b0 - start0, u1 = b1 - start1;
  if (u0 < 0) u0 += rows;
  if (u1 < 0) u1 += cols;
  const size_t index = (size_t)u1 * rows + u0;
  out[n - index - 1] = (int8_t)value;
}

/* MapGlobalPlanner::ifBlocked for a batch of query points (move_control/include/move_control/map_global_planner.h:
 * 39-54 over grid_map::CircleIterator, grid_map_core/src/iterators/CircleIterator.cpp): one warp per query, lanes
 * stride over the cells of the circle's bounding block; out[q] = 1 if any cell centre within `radius` holds a
 * non-NaN value > 0.  Lets the host-side RRT planner test candidates against the device-resident master layer
 * without downloading it (rrt_planner.cpp:53). */
__global__ void grid_blocked_kernel(GridDims d, RobotGeom g, const LayerRef layer,
                                    const double* __restrict__ xy, int n, double radius, uint8_t* __restrict__ out) {
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (q >= n) return;
  const double x = xy[2 * q], y = xy[2 * q + 1];
  const double r2 = radius * radius;
  double tlx = x + radius, tly = y + radius, brx = x - radius, bry = y - radius;
  limit_position_to_range(tlx, tly, d.len_x, d.len_y, g.pos_x, g.pos_y);
  limit_position_to_range(brx, bry, d.len_x, d.len_y, g.pos_x, g.pos_y);
  int sr, sc, er, ec;
  bool hit = false;
  if (grid_index(d, g, tlx, tly, sr, sc) && grid_index(d, g, brx, bry, er, ec)) {
    if ((g.start0 | g.start1) != 0) {
      sr -= g.start0;
      sc -= g.start1;
      er -= g.start0;
      ec -= g.start1;
      wrap_index(sr, d.rows);
      wrap_index(sc, d.cols);
      wrap_index(er, d.rows);
      wrap_index(ec, d.cols);
    }
    const int nr = er - sr + 1, nc = ec - sc + 1;
    for (int k = lane; k < nr * nc && nr > 0 && nc > 0; k += 32) {
      int b0 = sr + k % nr, b1 = sc + k / nr;
      if (b0 < 0 || b1 < 0 || b0 >= d.rows || b1 >= d.cols) continue;
      if ((g.start0 | g.start1) != 0) {
        b0 += g.start0;
        b1 += g.start1;
        wrap_index(b0, d.rows);
        wrap_index(b1, d.cols);
      }
      double px, py;
      position_from_index(b0, b1, d.len_x, d.len_y, g.pos_x, g.pos_y, d.res, d.rows, d.cols, g.start0, g.start1, px, py);
      const double dx = px - x, dy = py - y;
      if (!(dx * dx + dy * dy <= r2)) continue;
      const float v = layer.at(b0, b1);
      if (v > 0.0f) hit = true; /* false for NaN */
    }
  }
  const unsigned any = __ballot_sync(0xffffffffu, hit);
  if (lane == 0) out[q] = any ? 1 : 0;
}

/* The two-layer compose MapProvider::composeMasterMapFromLayerdMap carries commented out
 * (move_control/src/map_provider.cpp:218-220): master = (range is NaN and laser is not ? 0 : range) +
 * (laser is NaN and range is not ? 0 : laser), i.e. the sum where both are known, the known one where only one is,
 * NaN where neither is.  Sources in either layer format (one robot's LayerRef stride apart), destination float. */
__global__ void grid_compose_kernel(LayerRef a, size_t a_stride, LayerRef b, size_t b_stride, float* __restrict__ dst,
                                    int rows, int cols, size_t n_total) {
  const size_t per = (size_t)rows * cols;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_total; i += stride) {
    const size_t robot = i / per, lin = i - robot * per;
    const int r = (int)(lin % rows), c = (int)(lin / rows);
    LayerRef ar = a, br = b;
    ar.base = static_cast<const char*>(a.base) + robot * a_stride;
    br.base = static_cast<const char*>(b.base) + robot * b_stride;
    const float x = ar.at(r, c), y = br.at(r, c);
    const float x2 = (isnan(x) && !isnan(y)) ? 0.0f : x;
    const float y2 = (isnan(y) && !isnan(x)) ? 0.0f : y;
    dst[i] = x2 + y2;
  }
}

/* Measurement aid (bench.py): stream `n16` 16-byte words through L2 - write them (mode 1), or read them and fold
 * the result into *sink (mode 2) - so that nothing of the previous step is left in the cache. */
__global__ void l2_flush_kernel(uint4* __restrict__ buf, size_t n16, int mode, unsigned* __restrict__ sink) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  unsigned acc = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
    if (mode == 1) {
      buf[i] = make_uint4((unsigned)i, 0u, 0u, 0u);
    } else {
      const uint4 v = buf[i];
      acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
  }
  if (mode == 2 && acc == 0x9e3779b9u) *sink = acc; /* keeps the loads alive */
}

/* Measurement aid (bench.py): the rate at which this GPU retires UNORDERED 4-byte reductions (RED.ADD, no
 * return value) at uniformly random words of a buffer - the ceiling of the obvious "one atomic per cell visit" design
 * (which could not be exact anyway: clear and mark do not commute).  The tile kernel's visits per second are quoted
 * against it. */
__global__ void __launch_bounds__(256) red_calibration_kernel(unsigned* __restrict__ buf, unsigned word_mask, int iters,
                                                              unsigned seed) {
  unsigned x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + seed;
#pragma unroll 8
  for (int i = 0; i < iters; i++) {
    x ^= x << 13; /* xorshift32 */
    x ^= x >> 17;
    x ^= x << 5;
    atomicAdd(&buf[x & word_mask], 1u);
  }
}

}  // namespace b200nav
#endif

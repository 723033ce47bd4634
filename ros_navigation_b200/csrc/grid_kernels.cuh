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
__global__ void grid_to_occupancy_kernel(const float* __restrict__ layer, int rows, int cols, int start0, int start1,
                                         float data_min, float data_max, int8_t* __restrict__ out) {
  const size_t n = (size_t)rows * cols;
  const size_t lin = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (lin >= n) return;
  const int b0 = (int)(lin % rows), b1 = (int)(lin / rows);
  float value = (layer[lin] - data_min) / (data_max - data_min);
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

}  // namespace b200nav
#endif

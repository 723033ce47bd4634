/*
 * himm_kernels.cuh -- HIMM certainty-grid update on sm_100a.
 *
 * Replaces (behaviour, not code): MapUpdater::lineOnMap / clearCell / markCell
 * (move_control/include/move_control/map_updater.h:38-71), the in-order sample loop of
 * LaserMapUpdater::updateMap (move_control/src/laser_map_updater.cpp:15-20) and grid_map::LineIterator
 * (grid_map_core/src/iterators/LineIterator.cpp:16-150).
 *
 * Design (see DESIGN.md "HIMM"):
 *   K0 himm_prep_kernel   one thread per RangeSample: fp64 clip of both ray ends into the map and
 *                         position->index, producing a 24-byte BeamSeg (integer Bresenham end points + mark cell).
 *   K1 himm_tile_kernel   one CTA per (grid tile, robot).  A tile is split into SUB x SUB sub-tiles, each OWNED by
 *                         one warp and staged in shared memory.  The owning warp applies every beam that crosses its
 *                         sub-tile strictly in sample order (clear along the Bresenham cells, then the +30 mark),
 *                         32 lanes striding over the cells of ONE beam through a closed form of the Bresenham
 *                         recurrence.  Because a cell is only ever touched by its owner warp, in order, no atomics
 *                         are needed and the saturating clear/mark sequence is reproduced bit-exactly - the result
 *                         cannot depend on scheduling.  Only the 256-byte column segments a beam actually crosses
 *                         are loaded from / stored to HBM (coalesced, full sectors).
 */
#ifndef B200NAV_HIMM_KERNELS_CUH
#define B200NAV_HIMM_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200nav.h"
#include "geometry.h"

namespace b200nav {

struct HimmArgs {
  GridDims dims;
  const RobotGeom* geom;          /* [n_robots]                                   */
  float* layer;                   /* [n_robots][cols][rows]                       */
  const b200nav_sample* samples;  /* device                                       */
  const int32_t* offsets;         /* device [n_robots+1], or NULL in single mode  */
  BeamSeg* segs;                  /* device scratch [total]                       */
  int robot0;                     /* first robot handled by blockIdx.y == 0       */
  int n_active;                   /* robots handled by this launch                */
  int single_n;                   /* >= 0: single-robot mode, samples [0, n)      */
  int total;                      /* total samples                                */
  int tiles_r, tiles_c;           /* CTA tiles per grid                           */
};

/* ---------------------------------------------------------------------------------------------------------------
 * K0: RangeSample -> BeamSeg
 * ------------------------------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(128) himm_prep_kernel(HimmArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.total) return;
  int robot;
  if (a.single_n >= 0) {
    robot = a.robot0;
  } else {
    /* last r with offsets[r] <= i */
    int lo = 0, hi = a.n_active;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(&a.offsets[mid]) <= i) lo = mid;
      else hi = mid;
    }
    robot = a.robot0 + lo;
  }
  const RobotGeom g = a.geom[robot];
  const b200nav_sample s = a.samples[i];
  a.segs[i] = make_beam(a.dims, g, s.sx, s.sy, s.ex, s.ey, s.clear_end);
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
 * K1: tile-owner update kernel.
 *   SUB        sub-tile edge (cells) owned by one warp; smem pitch SUB+1 floats -> conflict-free along both axes
 *   WR x WC    warps per CTA tile (rows x cols) -> CTA tile = (SUB*WR) x (SUB*WC) cells
 *   LIST_CAP   beams per chunk (list of beams crossing the CTA tile, kept in sample order)
 * ------------------------------------------------------------------------------------------------------------- */
template <int SUB, int WR, int WC, int LIST_CAP>
struct HimmTileCfg {
  static constexpr int kWarps = WR * WC;
  static constexpr int kThreads = 32 * kWarps;
  static constexpr int kPitch = SUB + 1;
  static constexpr int kTileR = SUB * WR;
  static constexpr int kTileC = SUB * WC;
  static constexpr int kSubFloats = SUB * kPitch;
  static constexpr size_t kSmemBytes = sizeof(float) * kSubFloats * kWarps + sizeof(uint16_t) * LIST_CAP +
                                       sizeof(int) * (2 * kWarps);
  static constexpr int kColWords = SUB / 32; /* 32-bit words of the per-warp column mask */
  static_assert(SUB % 32 == 0 && SUB <= 128, "SUB must be 32, 64, 96 or 128");
};

template <int SUB>
struct ColMask {
  uint32_t w[SUB / 32];
};

template <int SUB>
__device__ __forceinline__ void colmask_add_range(ColMask<SUB>& m, int lo, int hi) {
#pragma unroll
  for (int k = 0; k < SUB / 32; k++) {
    const int a = max(lo - 32 * k, 0), b = min(hi - 32 * k, 31);
    if (a <= b) m.w[k] |= (0xffffffffu >> (31 - (b - a))) << a;
  }
}

template <int SUB, int WR, int WC, int LIST_CAP>
__global__ void __launch_bounds__(32 * WR * WC) himm_tile_kernel(HimmArgs a) {
  using Cfg = HimmTileCfg<SUB, WR, WC, LIST_CAP>;
  extern __shared__ __align__(16) unsigned char himm_smem_raw[];
  float* tiles = reinterpret_cast<float*>(himm_smem_raw);
  uint16_t* list = reinterpret_cast<uint16_t*>(tiles + Cfg::kSubFloats * Cfg::kWarps);
  int* warp_cnt = reinterpret_cast<int*>(list + LIST_CAP);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int robot = a.robot0 + blockIdx.y;
  const int tile_r = blockIdx.x % a.tiles_r, tile_c = blockIdx.x / a.tiles_r;
  const int rows = a.dims.rows, cols = a.dims.cols;

  /* CTA tile rectangle (inclusive, clipped to the grid). */
  const int TR0 = tile_r * Cfg::kTileR, TC0 = tile_c * Cfg::kTileC;
  const int TR1 = min(TR0 + Cfg::kTileR, rows) - 1, TC1 = min(TC0 + Cfg::kTileC, cols) - 1;
  /* This warp's sub-tile rectangle. */
  const int wr = warp % WR, wc = warp / WR;
  const int R0 = TR0 + wr * SUB, C0 = TC0 + wc * SUB;
  const int R1 = min(R0 + SUB, rows) - 1, C1 = min(C0 + SUB, cols) - 1;
  const bool warp_has_cells = (R0 <= R1) && (C0 <= C1);

  int beg, end;
  if (a.single_n >= 0) {
    beg = 0;
    end = a.single_n;
  } else {
    beg = __ldg(&a.offsets[blockIdx.y]);
    end = __ldg(&a.offsets[blockIdx.y + 1]);
  }
  if (beg >= end) return;

  float* tile = tiles + warp * Cfg::kSubFloats;
  float* gbase = a.layer + (size_t)robot * rows * cols;
  ColMask<SUB> loaded;
#pragma unroll
  for (int k = 0; k < SUB / 32; k++) loaded.w[k] = 0u;

  for (int base = beg; base < end; base += LIST_CAP) {
    const int chunk_end = min(base + LIST_CAP, end);

    /* ---- CTA filter: ordered list of the beams whose bounding box (or mark cell) touches the CTA tile ---- */
    int n_list = 0; /* uniform across the CTA */
    for (int i0 = base, it = 0; i0 < chunk_end; i0 += Cfg::kThreads, it++) {
      const int i = i0 + tid;
      bool hit = false;
      if (i < chunk_end) {
        const BeamSeg b = a.segs[i];
        if (b.r0 >= 0) {
          hit = max(b.r0, b.r1) >= TR0 && min(b.r0, b.r1) <= TR1 && max(b.c0, b.c1) >= TC0 && min(b.c0, b.c1) <= TC1;
        }
        if (b.mr >= 0) hit = hit || (b.mr >= TR0 && b.mr <= TR1 && b.mc >= TC0 && b.mc <= TC1);
      }
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      int* cnt = warp_cnt + (it & 1) * Cfg::kWarps; /* double-buffered: one barrier per step */
      if (lane == 0) cnt[warp] = __popc(bal);
      __syncthreads();
      int off = n_list, tot = 0;
#pragma unroll
      for (int w = 0; w < Cfg::kWarps; w++) {
        const int cw = cnt[w];
        if (w < warp) off += cw;
        tot += cw;
      }
      if (hit) list[off + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)(i - base);
      n_list += tot;
    }
    __syncthreads();

    if (warp_has_cells && n_list > 0) {
      /* ---- pass A: which columns of this warp's sub-tile will be touched? ---- */
      ColMask<SUB> need;
#pragma unroll
      for (int k = 0; k < SUB / 32; k++) need.w[k] = 0u;
      for (int j0 = 0; j0 < n_list; j0 += 32) {
        const int j = j0 + lane;
        if (j < n_list) {
          const BeamSeg b = a.segs[base + list[j]];
          if (b.r0 >= 0) {
            const LineForm f = line_form(b);
            int t0, t1;
            if (clip_line_to_rect(f, R0, R1, C0, C1, t0, t1)) {
              /* columns at t0 and t1 (monotone in between) */
              int ca, cb;
              if (f.row_major) {
                const int den = max(f.den, 1);
                ca = f.n0 + f.sn * (int)(((unsigned)(f.den >> 1) + (unsigned)t0 * (unsigned)f.add) / (unsigned)den);
                cb = f.n0 + f.sn * (int)(((unsigned)(f.den >> 1) + (unsigned)t1 * (unsigned)f.add) / (unsigned)den);
              } else {
                ca = f.m0 + f.sm * t0;
                cb = f.m0 + f.sm * t1;
              }
              colmask_add_range<SUB>(need, min(ca, cb) - C0, max(ca, cb) - C0);
            }
          }
          if (b.mr >= R0 && b.mr <= R1 && b.mc >= C0 && b.mc <= C1) colmask_add_range<SUB>(need, b.mc - C0, b.mc - C0);
        }
      }
#pragma unroll
      for (int k = 0; k < SUB / 32; k++) {
        need.w[k] = __reduce_or_sync(0xffffffffu, need.w[k]) & ~loaded.w[k];
      }
      /* ---- stage the newly needed columns (coalesced: one column = SUB consecutive floats) ---- */
#pragma unroll
      for (int k = 0; k < SUB / 32; k++) {
        unsigned m = need.w[k];
        loaded.w[k] |= m;
        while (m) {
          const int c = 32 * k + __ffs(m) - 1;
          m &= m - 1;
          const float* src = gbase + (size_t)(C0 + c) * rows + R0;
#pragma unroll
          for (int r = lane; r < SUB; r += 32)
            if (R0 + r <= R1) tile[c * Cfg::kPitch + r] = __ldg(src + r);
        }
      }
      __syncwarp();

      /* ---- pass B: apply the beams in sample order ---- */
      for (int j0 = 0; j0 < n_list; j0 += 32) {
        const int j = j0 + lane;
        /* lane-parallel set-up of up to 32 segments */
        int my_len = 0, my_off0 = 0, my_rem0 = 0, my_dm = 0, my_dn = 0, my_add = 0, my_den = 1, my_moff = -1;
        if (j < n_list) {
          const BeamSeg b = a.segs[base + list[j]];
          if (b.r0 >= 0) {
            const LineForm f = line_form(b);
            int t0, t1;
            if (clip_line_to_rect(f, R0, R1, C0, C1, t0, t1)) {
              const unsigned den = (unsigned)max(f.den, 1);
              const unsigned x0 = (unsigned)(f.den >> 1) + (unsigned)t0 * (unsigned)f.add;
              const unsigned q0 = x0 / den;
              my_rem0 = (int)(x0 - q0 * den);
              const int mj = f.m0 + f.sm * t0, mn = f.n0 + f.sn * (int)q0;
              const int r = f.row_major ? mj : mn, c = f.row_major ? mn : mj;
              my_off0 = (c - C0) * Cfg::kPitch + (r - R0);
              my_dm = f.row_major ? f.sm : f.sm * Cfg::kPitch;
              my_dn = f.row_major ? f.sn * Cfg::kPitch : f.sn;
              my_add = f.add;
              my_den = (int)den;
              my_len = t1 - t0 + 1;
            }
          }
          if (b.mr >= R0 && b.mr <= R1 && b.mc >= C0 && b.mc <= C1)
            my_moff = (b.mc - C0) * Cfg::kPitch + (b.mr - R0);
        }
        unsigned active = __ballot_sync(0xffffffffu, my_len > 0 || my_moff >= 0);
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
            /* lane L starts at step t0+L: floor((rem0 + L*add)/den) <= 32, exact through a float reciprocal
             * (x + 0.5 keeps the quotient >= 0.5/den away from an integer; float error here < 5e-6). */
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
              tile[off] = himm_clear(tile[off]);
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
            if (lane == 0) tile[moff] = himm_mark(tile[moff]);
            __syncwarp();
          }
        }
      }
    }
    __syncthreads(); /* the list is rewritten by the next chunk */
  }

  /* ---- write back the staged (== touched) columns ---- */
  if (warp_has_cells) {
#pragma unroll
    for (int k = 0; k < SUB / 32; k++) {
      unsigned m = loaded.w[k];
      while (m) {
        const int c = 32 * k + __ffs(m) - 1;
        m &= m - 1;
        float* dst = gbase + (size_t)(C0 + c) * rows + R0;
#pragma unroll
        for (int r = lane; r < SUB; r += 32)
          if (R0 + r <= R1) dst[r] = tile[c * Cfg::kPitch + r];
      }
    }
  }
}

}  // namespace b200nav
#endif

/*
 * vfh_kernels.cuh -- VFH+ steering decision on sm_100a, one CTA per robot, all stages fused.
 *
 * Replaces (behaviour, not code):
 *   Steerer::getRangesFromSubmap            move_control/src/steerer.cpp:147-191   (stage R)
 *   VFH::Update_VFH and its callees         move_control/src/vfh.cpp:480-605
 *     Calculate_Cells_Mag                   :986-1049   (stage M)
 *     Build_Primary_Polar_Histogram         :1057-1095  (stage H)
 *     Build_Binary_Polar_Histogram          :1102-1121  (stage B)
 *     Build_Masked_Polar_Histogram          :1131-1213  (stage K)
 *     Select_Direction / Select_Candidate_Angle / Cant_Turn_To_Goal / Set_Motion
 *                                           :755-870, :715-749, :612-654, :1222-1261 (stage S)
 *
 * Numerics: all float/double expressions keep the reference's C++ promotions and evaluation order and the file is
 * compiled with --fmad=false.  The primary histogram is summed by one thread per sector over the occupied cells in
 * the reference's (y outer, x inner) order, so it is bit-identical, not just within tolerance.  phi_left/phi_right
 * are order-free max/min and use a warp-shuffle + shared-memory reduction.
 * The active window is staged into shared memory by TMA (cp.async.bulk.tensor) when the layer pitch allows it
 * (rows % 4 == 0 and the window does not wrap the circular buffer), otherwise by coalesced loads.
 */
#ifndef B200NAV_VFH_KERNELS_CUH
#define B200NAV_VFH_KERNELS_CUH

#include <cuda.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/b200nav.h"
#include "cells.cuh"
#include "geometry.h"
#include "vfh_tables.h"

namespace b200nav {

struct VfhRobotState {
  float picked, last_picked, desired, blocked_radius;
  int last_chosen_speed, max_speed_for_picked;
};

struct VfhDev {
  VfhConst c;
  int nf; /* front cells = front_rows * window */
  const float* dir;
  const float* dist;
  const float* base;
  const double* thr;
  const int16_t* kidx;
  const uint32_t* masks; /* [table][f][nwords] */
  const int32_t* mtr;    /* Min_Turning_Radius[0..current_max_speed] */
  /* per-robot state */
  float* origin_hist;    /* [n][H] */
  float* hist;           /* [n][H] */
  float* last_binary;    /* [n][H] */
  VfhRobotState* st;     /* [n]    */
  double* ranges;        /* [n][361] pseudo-scan of the last update */
};

struct VfhGridArgs {
  GridDims dims;
  const RobotGeom* geom;
  const void* layer; /* FLOAT: float [n_robots][cols][rows]; CODED: bytes [n_robots][tiles][4352] (cells.cuh) */
  int coded;
  int tiles_r, tiles_c;
  int box_r, box_c; /* TMA box (floats) = smem window pitch / columns */
  int use_tma;
};

/* Fused command exchange ("peer push"): in the batched multi-GPU mode the kernel that takes the steering decision
 * also delivers it - thread 0 of every block stores the robot's 16-byte command straight into row (row0 + robot) of
 * the command table of EVERY rank through NVLink peer mappings, overlapped with the other blocks' computation.
 * fleet_flag_kernel (stream-ordered after the update, on the fleet's own stream) then publishes the cycle's epoch in
 * every rank's flag array; the consumer side is fleet_wait_kernel.  world == 0: no exchange. */
#define B200NAV_MAX_PEERS 16
struct VfhPush {
  b200nav_command* tables[B200NAV_MAX_PEERS]; /* table of rank p for this slot (peer-mapped)      */
  unsigned long long* flags[B200NAV_MAX_PEERS]; /* flag array of rank p for this slot: [world]     */
  int world, rank, row0;
  unsigned long long epoch;
};

/* Runs after the VFH+ kernel of the cycle has completed (all its peer stores performed): tells every rank that this
 * rank's rows of the slot are in place. */
__global__ void fleet_flag_kernel(const VfhPush push) {
  const int p = threadIdx.x;
  __threadfence_system();
  if (p < push.world) *reinterpret_cast<volatile unsigned long long*>(&push.flags[p][push.rank]) = push.epoch;
}

/* Flow control of the peer push.  A writer may only store cycle k+2's rows into a peer's table of slot s after that
 * peer has finished READING the cycle-k table of slot s (publishing cycle k is not enough: a rank that runs ahead
 * would tear the table under a slower reader).  Every rank therefore acknowledges, in every WRITER's ack array, the
 * last epoch of the slot it has consumed (fleet_release_kernel, stream-ordered after its reads), and a writer waits
 * for all acknowledgements of the previous epoch before its VFH+ kernel runs (fleet_wait_acks_kernel). */
struct FleetAck {
  unsigned long long* acks[B200NAV_MAX_PEERS]; /* ack array of rank p for this slot: [world], peer-mapped */
  int world, rank;
  unsigned long long epoch;
};
__global__ void fleet_release_kernel(const FleetAck a) {
  const int p = threadIdx.x;
  __threadfence_system();
  if (p < a.world) *reinterpret_cast<volatile unsigned long long*>(&a.acks[p][a.rank]) = a.epoch;
}
__global__ void fleet_wait_acks_kernel(const unsigned long long* acks, int world, unsigned long long epoch, int* err) {
  const int r = threadIdx.x;
  if (r < world) {
    const volatile unsigned long long* f = acks + r;
    long long spins = 0;
    while (*f < epoch) {
      __nanosleep(200);
      if (++spins > 5000000ll) { /* about a second: a peer died or never released the slot */
        *err = 2;
        break;
      }
    }
  }
  __threadfence_system();
}

#define B200NAV_VFH_THREADS 128
#define B200NAV_VFH_MAX_SECTORS 384 /* 360 / sector_angle, sector_angle >= 1 */
#define B200NAV_NRANGES 361

/* ---- reference helper expressions (same promotions as vfh.cpp) ------------------------------------------------ */
__device__ __forceinline__ int vfh_max_turnrate(const VfhConst& c, int speed) {
  int val = (c.max_turnrate_0ms - (int)(speed * (c.max_turnrate_0ms - c.max_turnrate_1ms) / 1000.0));
  return val < 0 ? 0 : val;
}
__device__ __forceinline__ int vfh_safety_dist(const VfhConst& c, int speed) {
  int val = (int)(c.safety_dist_0ms + (int)(speed * (c.safety_dist_1ms - c.safety_dist_0ms) / 1000.0));
  return val < 0 ? 0 : val;
}
__device__ __forceinline__ float vfh_bin_low(const VfhConst& c, int speed) {
  return (float)(c.bin_low_0ms - (speed * (c.bin_low_0ms - c.bin_low_1ms) / 1000.0));
}
__device__ __forceinline__ float vfh_bin_high(const VfhConst& c, int speed) {
  return (float)(c.bin_high_0ms - (speed * (c.bin_high_0ms - c.bin_high_1ms) / 1000.0));
}
__device__ __forceinline__ int vfh_speed_index(const VfhConst& c, int speed) {
  int val = (int)floorf(((float)speed / (float)c.current_max_speed) * c.num_tables);
  return val >= c.num_tables ? c.num_tables - 1 : val;
}
__device__ __forceinline__ float vfh_delta_angle(float a1, float a2) {
  float diff = a2 - a1;
  if (diff > 180) diff -= 360;
  else if (diff < -180) diff += 360;
  return diff;
}
/* glibc hypotf: (float)sqrt((double)x*x + (double)y*y) - products exact, one rounding in the sum, IEEE sqrt. */
__device__ __forceinline__ float vfh_hypotf(float x, float y) {
  return (float)sqrt((double)x * (double)x + (double)y * (double)y);
}

/* fmod(fmod(a, p) + p, p) (steerer.cpp:165-168) without the slow fp64 fmod in the common range: fmod is exact, and so
 * is a - p for p <= a <= 2p (Sterbenz), so both agree bit for bit; anything else (huge or non-finite angles, the sum
 * rounding up to 2p) takes the library call. */
__device__ __forceinline__ double vfh_wrap_two_pi(double a, double p) {
  double m = a;
  if (!(fabs(a) < p)) {
    if (fabs(a) < 2.0 * p) m = (a > 0) ? a - p : a + p;
    else m = fmod(a, p);
  }
  const double b = m + p;
  if (b < p) return b;
  if (b < 2.0 * p) return b - p;
  return fmod(b, p);
}

/* Positive doubles order like their bit patterns: shared-memory atomicMin on the pattern is an exact min. */
__device__ __forceinline__ void atomic_min_pos_double(double* addr, double v) {
  atomicMin(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

#ifdef VFH_STAGE_CLOCKS
__device__ unsigned long long g_vfh_clk[8];
#define VFH_CLK(k)                                                        \
  do {                                                                    \
    if (threadIdx.x == 0) {                                               \
      const long long now_ = clock64();                                   \
      atomicAdd(&g_vfh_clk[k], (unsigned long long)(now_ - clk_prev_));   \
      clk_prev_ = now_;                                                   \
    }                                                                     \
  } while (0)
#else
#define VFH_CLK(k)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

/* ----------------------------------------------------------------------------------------------------------------
 * Fused kernel.  FROM_GRID: build the pseudo-scan from the grid window (stage R), else take dev_ranges.
 * Dynamic shared memory layout (bytes): ranges[361] double | nz[nf] u16 | hist[H] float | window floats (TMA).
 * -------------------------------------------------------------------------------------------------------------- */
/* MIN_BLOCKS: resident CTAs per SM the register allocation must allow.  1 (70 registers) is fastest while a launch
 * is a single wave and the block's dependent chain sets the time (1024 robots: 23 vs 27 us); 16 (32 registers, a few
 * spills) wins once there are many waves (16 384 robots: 213 vs 272 us).  The host picks by robot count. */
template <bool FROM_GRID, int MIN_BLOCKS>
__global__ void __launch_bounds__(B200NAV_VFH_THREADS, MIN_BLOCKS)
vfh_update_kernel(const VfhDev v, const VfhGridArgs ga, const __grid_constant__ CUtensorMap tmap,
                  const b200nav_vfh_input* __restrict__ in, const double* __restrict__ dev_ranges,
                  b200nav_command* __restrict__ out, int robot0, const VfhPush push) {
  extern __shared__ __align__(128) unsigned char vfh_smem_raw[];
  const VfhConst& c = v.c;
  const int H = c.hist_size, W = c.window, nf = v.nf;
  /* carve */
  float* s_window = reinterpret_cast<float*>(vfh_smem_raw); /* 128B aligned, TMA destination (FROM_GRID only) */
  size_t off = FROM_GRID ? (((size_t)ga.box_r * ga.box_c * sizeof(float) + 127) & ~(size_t)127) : 0;
  double* s_ranges = reinterpret_cast<double*>(vfh_smem_raw + off);
  off += sizeof(double) * (B200NAV_NRANGES + 1);
  float* s_hist = reinterpret_cast<float*>(vfh_smem_raw + off);
  off += sizeof(float) * ((H + 1) & ~1);
  uint16_t* s_nz = reinterpret_cast<uint16_t*>(vfh_smem_raw + off);

  __shared__ SubmapInfo s_sub;
  __shared__ int s_sub_ok;
  __shared__ int s_warp_cnt[4][B200NAV_VFH_THREADS / 32];
  __shared__ float s_red[2][B200NAV_VFH_THREADS / 32];
  __shared__ __align__(8) unsigned long long s_mbar;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef VFH_STAGE_CLOCKS
  long long clk_prev_ = clock64();
#endif
  const int robot = robot0 + blockIdx.x;
  const b200nav_vfh_input inp = in[blockIdx.x];
  uint32_t flags = 0;

  /* Cant_Turn_To_Goal's goal vector (vfh.cpp:612-654) only depends on the inputs: a lane of the last warp computes it
   * now (two fp64 sin/cos) instead of the selection thread at the very end of the chain */
  __shared__ float s_goal[2];
  __shared__ uint32_t s_blocked[B200NAV_VFH_MAX_SECTORS / 32 + B200NAV_VFH_THREADS / 32]; /* masked histogram == 1 */
  if (tid == B200NAV_VFH_THREADS - 32) {
    s_goal[0] = (float)(inp.goal_distance * cos((inp.goal_direction) * 3.14159265358979323846 / 180));
    s_goal[1] = (float)(inp.goal_distance * sin((inp.goal_direction) * 3.14159265358979323846 / 180));
  }

  /* ================= stage R: pseudo-scan ================= */
  if (FROM_GRID) {
    const RobotGeom g = ga.geom[robot];
    if (tid == 0) {
      SubmapInfo si;
      const bool ok = submap_info(ga.dims, g, inp.x, inp.y, c.submap_length, c.submap_length, si);
      s_sub = si;
      s_sub_ok = ok ? 1 : 0;
      /* TMA path needs a window that does not cross the circular-buffer seam */
      /* TMA (tiled, no interleave) needs the innermost coordinate 16-byte aligned: start the box at row
       * tl_r & ~3 and skip the first (tl_r & 3) floats of every staged column. */
      int tma = ga.use_tma && ok && si.size_r + (si.tl_r & 3) <= ga.box_r && si.size_c <= ga.box_c &&
                si.tl_r + si.size_r <= ga.dims.rows && si.tl_c + si.size_c <= ga.dims.cols;
      s_sub_ok |= tma ? 2 : 0;
      if (tma) {
        const uint32_t mbar = smem_u32(&s_mbar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t bytes = (uint32_t)(ga.box_r * ga.box_c * sizeof(float));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
            ::"r"(smem_u32(s_window)), "l"(&tmap), "r"(si.tl_r & ~3), "r"(si.tl_c), "r"(robot), "r"(mbar)
            : "memory");
      }
    }
    for (int i = tid; i < B200NAV_NRANGES; i += blockDim.x) s_ranges[i] = 5000.0;
    __syncthreads();
    VFH_CLK(0); /* submap info + TMA issue */
    const int sub_ok = s_sub_ok;
    if (sub_ok & 1) {
      const SubmapInfo si = s_sub;
      const bool tma = (sub_ok & 2) != 0;
      if (tma) { /* wait for the window */
        const uint32_t mbar = smem_u32(&s_mbar);
        uint32_t done = 0;
        while (!done) {
          asm volatile(
              "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
              : "=r"(done)
              : "r"(mbar), "r"(0)
              : "memory");
        }
      }
      LayerRef lay;
      lay.coded = ga.coded;
      lay.rows = ga.dims.rows;
      lay.tiles_r = ga.tiles_r;
      lay.base = ga.coded ? static_cast<const void*>(static_cast<const uint8_t*>(ga.layer) +
                                                     (size_t)robot * ga.tiles_r * ga.tiles_c * HIMM_TILE_BYTES)
                          : static_cast<const void*>(static_cast<const float*>(ga.layer) +
                                                     (size_t)robot * ga.dims.rows * ga.dims.cols);
      const int ncell = si.size_r * si.size_c;
      const double res = ga.dims.res;
      const double offx = si.pos_x + (0.5 * si.len_x - 0.5 * res), offy = si.pos_y + (0.5 * si.len_y - 0.5 * res);
      /* lin -> (i0, i1) = (lin % size_r, lin / size_r) without a division: ncell <= 2^17, size_r <= 2^8 */
      const unsigned inv_r = si.size_r > 1 ? (unsigned)((0x100000000ull + (unsigned)si.size_r - 1u) / (unsigned)si.size_r) : 0u;
      constexpr int kBatch = 8; /* window cells in flight per thread: one L2 round trip per batch, not per cell */
      for (int base = 0; base < ncell; base += kBatch * (int)blockDim.x) {
        float vals[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; u++) {
          const int lin = base + u * (int)blockDim.x + tid;
          vals[u] = __int_as_float(0x7fc00000);
          if (lin < ncell) {
            const int i1 = si.size_r > 1 ? (int)__umulhi((unsigned)lin, inv_r) : lin;
            const int i0 = lin - i1 * si.size_r;
            if (tma) {
              vals[u] = s_window[i1 * ga.box_r + i0 + (si.tl_r & 3)];
            } else {
              int b0 = si.utl_r + i0, b1 = si.utl_c + i1;
              if ((g.start0 | g.start1) != 0) {
                b0 += g.start0;
                b1 += g.start1;
                wrap_index(b0, ga.dims.rows);
                wrap_index(b1, ga.dims.cols);
              }
              vals[u] = lay.at(b0, b1);
            }
          }
        }
        /* occupied cells of this thread as a bit set: the expensive fp64 part below then runs once per occupied cell
         * of the busiest lane, not once per batch slot that is occupied in ANY lane of the warp */
        unsigned todo = 0u;
#pragma unroll
        for (int u = 0; u < kBatch; u++)
          if (!(isnan(vals[u]) || vals[u] <= c.occupied_threshold)) todo |= 1u << u;
        while (todo) {
          const int u = __ffs(todo) - 1;
          todo &= todo - 1u;
          const int lin = base + u * (int)blockDim.x + tid;
          const int i1 = si.size_r > 1 ? (int)__umulhi((unsigned)lin, inv_r) : lin;
          const int i0 = lin - i1 * si.size_r;
          const double px = offx + res * int_to_f64(-i0), py = offy + res * int_to_f64(-i1);
          const double angle = atan2(py - inp.y, px - inp.x);
          const double a = angle - inp.yaw + 3.14 / 2;
          const double twopi = 2.0 * 3.14159265358979323846;
          const double np = vfh_wrap_two_pi(a, twopi); /* == fmod(fmod(a, twopi) + twopi, twopi), bit for bit */
          const double deg = np * 180.0 / 3.14159265358979323846;
          if (!(deg <= 180)) continue; /* also rejects NaN poses */
          const int fl = (int)floor(deg), ce = (int)ceil(deg);
          const double dx = inp.x - px, dy = inp.y - py;
          const double distance = sqrt(dx * dx + dy * dy) * 1000.0;
          atomic_min_pos_double(&s_ranges[fl * 2], distance);
          atomic_min_pos_double(&s_ranges[ce * 2], distance);
        }
      }
    } else {
      flags |= B200NAV_CMD_NO_SUBMAP;
    }
    __syncthreads();
    for (int i = tid; i < B200NAV_NRANGES; i += blockDim.x) v.ranges[(size_t)robot * B200NAV_NRANGES + i] = s_ranges[i];
    VFH_CLK(1); /* window -> ranges */
  } else {
    for (int i = tid; i < B200NAV_NRANGES; i += blockDim.x) {
      const double r = dev_ranges[(size_t)blockIdx.x * 2 * B200NAV_NRANGES + 2 * i];
      s_ranges[i] = r;
      v.ranges[(size_t)robot * B200NAV_NRANGES + i] = r;
    }
    __syncthreads();
  }

  /* ================= Update_VFH prologue (vfh.cpp:491-515) ================= */
  VfhRobotState st = v.st[robot];
  st.desired = inp.goal_direction;
  int speed = inp.current_speed < 0 ? 0 : inp.current_speed;
  if (speed < st.last_chosen_speed) speed = st.last_chosen_speed;

  /* ================= stage M: cell magnitudes -> ordered list of occupied cells ================= */
  const float r_safe = c.robot_radius + (float)vfh_safety_dist(c, speed);
  int nz_count = 0;
  int emergency_local = 0;
  /* kSub sub-batches of blockDim cells per pass: their table loads are issued together and one block barrier
   * orders the whole pass (cell order f is preserved: sub-batch after sub-batch, warp after warp, lane after lane) */
  constexpr int kSub = 4;
  for (int f0 = 0; f0 < nf; f0 += kSub * (int)blockDim.x) {
    int kk[kSub];
#pragma unroll
    for (int u = 0; u < kSub; u++) {
      const int f = f0 + u * (int)blockDim.x + tid;
      kk[u] = (f < nf) ? (int)v.kidx[f] : -1;
    }
    unsigned bal[kSub];
    bool occ[kSub];
#pragma unroll
    for (int u = 0; u < kSub; u++) {
      const int f = f0 + u * (int)blockDim.x + tid;
      occ[u] = false;
      if (kk[u] >= 0 && v.thr[f] > s_ranges[kk[u]]) {
        const int x = f % W, y = f / W;
        if (v.dist[f] < r_safe && !(x == c.center && y == c.center)) emergency_local = 1;
        occ[u] = true;
      }
      bal[u] = __ballot_sync(0xffffffffu, occ[u]);
      if (lane == 0) s_warp_cnt[u][warp] = __popc(bal[u]);
    }
    __syncthreads();
    int base_pos = nz_count;
#pragma unroll
    for (int u = 0; u < kSub; u++) {
      int pos = base_pos, tot = 0;
#pragma unroll
      for (int w = 0; w < B200NAV_VFH_THREADS / 32; w++) {
        const int cw = s_warp_cnt[u][w];
        if (w < warp) pos += cw;
        tot += cw;
      }
      if (occ[u]) s_nz[pos + __popc(bal[u] & ((1u << lane) - 1u))] = (uint16_t)(f0 + u * (int)blockDim.x + tid);
      base_pos += tot;
    }
    nz_count = base_pos;
    __syncthreads(); /* s_warp_cnt is rewritten by the next pass */
  }
  const int emergency = __syncthreads_or(emergency_local);
  VFH_CLK(2); /* stage M */

  float* g_origin = v.origin_hist + (size_t)robot * H;
  float* g_hist = v.hist + (size_t)robot * H;
  float* g_last = v.last_binary + (size_t)robot * H;

  float phi_right = 0.f, phi_left = 180.f;
  if (emergency) {
    /* vfh.cpp:1070-1077: OriginHist all 1; Hist untouched */
    for (int s = tid; s < H; s += blockDim.x) g_origin[s] = 1.0f;
    flags |= B200NAV_CMD_EMERGENCY;
  } else {
    /* ================= stage H + B ================= */
    const int tab = vfh_speed_index(c, speed);
    const uint32_t* masks = v.masks + (size_t)tab * nf * c.nwords;
    const float high = vfh_bin_high(c, speed), low = vfh_bin_low(c, speed);
    for (int s = tid; s < H; s += blockDim.x) {
      float h = 0.f;
      const int word = s >> 5;
      const uint32_t bit = 1u << (s & 31);
      /* occupied cells in (y outer, x inner) order; the table loads of four cells are in flight together, the sum
       * keeps the reference's order */
      int j = 0;
      for (; j + 4 <= nz_count; j += 4) {
        uint32_t mw[4];
        float bv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int f = s_nz[j + u];
          mw[u] = __ldg(&masks[(size_t)f * c.nwords + word]);
          bv[u] = __ldg(&v.base[f]);
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (mw[u] & bit) h += bv[u];
      }
      for (; j < nz_count; j++) {
        const int f = s_nz[j];
        if (__ldg(&masks[(size_t)f * c.nwords + word]) & bit) h += __ldg(&v.base[f]);
      }
      g_origin[s] = h;
      float b;
      if (h > high) b = 1.0f;
      else if (h < low) b = 0.0f;
      else b = g_last[s];
      g_last[s] = b;
      s_hist[s] = b;
    }
    /* ================= stage K: blocked circles -> phi_right / phi_left ================= */
    int mtr_idx = speed;
    if (mtr_idx > c.current_max_speed) mtr_idx = c.current_max_speed; /* reference reads out of bounds here (H4c) */
    const int mtr = v.mtr[mtr_idx];
    const float cxr = c.center + (mtr / (float)c.cell_width);
    const float cxl = c.center - (mtr / (float)c.cell_width);
    const float cy = c.center;
    const float rb = mtr + c.robot_radius + vfh_safety_dist(c, speed);
    st.blocked_radius = rb;
    float pr = 0.f, pl = 180.f;
    for (int j = tid; j < nz_count; j += blockDim.x) {
      const int f = s_nz[j];
      const float d = v.dir[f];
      const int x = f % W, y = f / W;
      if (vfh_delta_angle(d, 90.f) > 0) {
        const float dist_r = vfh_hypotf(cxr - x, cy - y) * c.cell_width;
        if (dist_r < rb) pr = fmaxf(pr, d);
      } else {
        const float dist_l = vfh_hypotf(cxl - x, cy - y) * c.cell_width;
        if (dist_l < rb) pl = fminf(pl, d);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      pr = fmaxf(pr, __shfl_xor_sync(0xffffffffu, pr, o));
      pl = fminf(pl, __shfl_xor_sync(0xffffffffu, pl, o));
    }
    if (lane == 0) {
      s_red[0][warp] = pr;
      s_red[1][warp] = pl;
    }
    __syncthreads(); /* also publishes s_hist */
    VFH_CLK(3); /* stages H, B, K */
#pragma unroll
    for (int w = 0; w < B200NAV_VFH_THREADS / 32; w++) {
      phi_right = fmaxf(phi_right, s_red[0][w]);
      phi_left = fminf(phi_left, s_red[1][w]);
    }
    /* mask (vfh.cpp:1196-1210); the blocked sectors also go into a bit set for the selection stage */
    for (int s0 = 0; s0 < H; s0 += blockDim.x) {
      const int s = s0 + tid;
      bool blocked = false;
      if (s < H) {
        const float angle = (float)(s * c.sector_angle);
        blocked = !((s_hist[s] == 0) &&
                    (((vfh_delta_angle(angle, phi_right) <= 0) && (vfh_delta_angle(angle, 90.f) >= 0)) ||
                     ((vfh_delta_angle(angle, phi_left) >= 0) && (vfh_delta_angle(angle, 90.f) <= 0))));
        g_hist[s] = blocked ? 1.f : 0.f;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, blocked);
      if (lane == 0) s_blocked[(s0 >> 5) + warp] = bal;
    }
    __syncthreads();
    VFH_CLK(4); /* mask */
  }

  /* ================= stage S: direction, speed, turn rate (single thread) ================= */
  if (tid == 0) {
    if (emergency) {
      st.picked = st.last_picked;
      st.max_speed_for_picked = 0;
      st.last_picked = st.picked;
    } else {
      auto is_blocked = [&](int s) { return (s_blocked[s >> 5] >> (s & 31)) & 1u; };
      int start = -1;
      for (int i = 0; i < H / 2; i++)
        if (is_blocked(i)) {
          start = i;
          break;
        }
      if (start == -1) {
        st.picked = st.desired;
        st.last_picked = st.picked;
        st.max_speed_for_picked = c.current_max_speed;
      } else {
        /* openings -> candidates -> first minimum of the weight (vfh.cpp:786-868, 715-749) */
        int n_cand = 0;
        float best_angle = 90.f, min_weight = 10000000.f;
        int best_speed = st.max_speed_for_picked;
        auto consider = [&](float cand, int cand_speed) {
          const float weight = c.u1 * fabsf(vfh_delta_angle(st.desired, cand)) +
                               c.u2 * fabsf(vfh_delta_angle(st.last_picked, cand));
          if (weight < min_weight) {
            min_weight = weight;
            best_angle = cand;
            best_speed = cand_speed;
          }
          n_cand++;
        };
        const int cmax = c.current_max_speed;
        const int sp_narrow = (cmax < c.max_speed_narrow) ? cmax : c.max_speed_narrow;
        const int sp_wide = (cmax < c.max_speed_wide) ? cmax : c.max_speed_wide;
        /* The reference walks the H + 1 positions n = 0..H, sector s(n) = (start + n) % H, and toggles between
         * "looking for a free sector" (-> first) and "looking for a blocked one" (-> second), vfh.cpp:786-868.  Only
         * those events matter, so they are found with bit scans over the blocked-sector set instead of visiting every
         * position.  s(0) = s(H) = start is blocked, so every opening is closed by n = H at the latest. */
        auto find = [&](int lo, int hi, bool want) -> int { /* first s in [lo, hi) with blocked == want, else -1 */
          while (lo < hi) {
            const int w = lo >> 5, lim = min(hi, (w + 1) << 5);
            uint32_t bits = want ? s_blocked[w] : ~s_blocked[w];
            bits &= 0xffffffffu << (lo & 31);
            if (lim - (w << 5) < 32) bits &= (1u << (lim - (w << 5))) - 1u;
            if (bits) return (w << 5) + __ffs(bits) - 1;
            lo = lim;
          }
          return -1;
        };
        const int wrap_at = H - start; /* positions n >= wrap_at map to s = n - wrap_at */
        auto next_event = [&](int n, bool want) -> int { /* first n' in [n, H] with blocked(s(n')) == want, else H+1 */
          if (n < wrap_at) {
            const int s = find(start + n, H, want);
            if (s >= 0) return s - start;
            n = wrap_at;
          }
          if (n <= H) {
            const int s = find(n - wrap_at, start + 1, want);
            if (s >= 0) return s + wrap_at;
          }
          return H + 1;
        };
        int first = 0, second = 0;
        for (int n = 0;;) {
          n = next_event(n, false); /* free sector while looking left */
          if (n > H) break;
          first = ((n < wrap_at) ? start + n : n - wrap_at) * c.sector_angle;
          n = next_event(n + 1, true); /* the blocked sector that closes the opening */
          if (n > H) break;
          const int s = (n < wrap_at) ? start + n : n - wrap_at;
          n++;
          second = (s - 1) * c.sector_angle;
          if (second < 0) second += 360;
          const float angle = vfh_delta_angle((float)first, (float)second);
          if (fabsf(angle) < 10) continue;
          const float centre = (float)(first + (second - first) / 2.0);
          if (fabsf(angle) < 80) {
            consider(centre, sp_narrow);
          } else {
            consider(centre, cmax);
            const float c2 = (float)((first + 40) % 360);
            consider(c2, sp_wide);
            float c3 = (float)(second - 40);
            if (c3 < 0) c3 += 360;
            consider(c3, sp_wide);
            if ((vfh_delta_angle(st.desired, c2) < 0) && (vfh_delta_angle(st.desired, c3) > 0))
              consider(st.desired, sp_wide);
          }
        }
        if (n_cand == 0) {
          st.picked = st.last_picked;
          st.max_speed_for_picked = 0;
          st.last_picked = st.picked;
          flags |= B200NAV_CMD_HEMMED_IN;
        } else {
          st.picked = best_angle;
          st.max_speed_for_picked = best_speed;
          st.last_picked = st.picked;
        }
      }
    }
    /* speed (vfh.cpp:566-599) */
    int speed_incr;
    if ((inp.dt > 0.3) || (inp.dt < 0)) speed_incr = 10;
    else speed_incr = (int)(c.max_acceleration * inp.dt);
    {
      /* Cant_Turn_To_Goal (vfh.cpp:612-654) */
      const float goal_x = s_goal[0], goal_y = s_goal[1]; /* st.desired == inp.goal_direction */
      const float rb = st.blocked_radius;
      bool cant = false;
      float dc = vfh_hypotf(goal_x - rb, goal_y);
      if (dc + inp.goal_tolerance < rb) cant = true;
      if (!cant) {
        dc = vfh_hypotf(-goal_x - rb, goal_y);
        if (dc + inp.goal_tolerance < rb) cant = true;
      }
      if (cant) {
        speed_incr = -speed_incr;
        flags |= B200NAV_CMD_CANT_TURN;
      }
    }
    int chosen_speed = st.last_chosen_speed + speed_incr;
    if (!(chosen_speed < st.max_speed_for_picked)) chosen_speed = st.max_speed_for_picked;
    /* Set_Motion (vfh.cpp:1222-1261) */
    int turnrate;
    const int mt = vfh_max_turnrate(c, speed);
    if (chosen_speed <= 0) {
      turnrate = mt;
      chosen_speed = 0;
    } else {
      if ((st.picked > 270) && (st.picked < 360)) {
        turnrate = -1 * mt;
      } else if ((st.picked < 270) && (st.picked > 180)) {
        turnrate = mt;
      } else {
        turnrate = (int)rint(((float)(st.picked - 90) / 75.0) * mt);
        if (turnrate > mt) turnrate = mt;
        else if (turnrate < (-1 * mt)) turnrate = -1 * mt;
      }
    }
    st.last_chosen_speed = chosen_speed;
    v.st[robot] = st;
    b200nav_command cmd;
    cmd.speed = chosen_speed;
    cmd.turnrate = turnrate;
    cmd.picked_angle = st.picked;
    cmd.flags = flags;
    out[blockIdx.x] = cmd;
    /* fused exchange (batched multi-GPU mode): the decision goes straight into row (row0 + robot) of every rank's
     * command table through the NVLink peer mappings - plain stores, no fence: the epoch that tells the consumers the
     * rows are complete is published by fleet_flag_kernel, stream-ordered after this kernel */
    for (int p = 0; p < push.world; p++)
      asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(&push.tables[p][push.row0 + blockIdx.x]),
                   "r"((unsigned)cmd.speed), "r"((unsigned)cmd.turnrate), "r"(__float_as_uint(cmd.picked_angle)),
                   "r"(cmd.flags)
                   : "memory");
    VFH_CLK(5); /* stage S */
  }

}

/* Consumer side of the peer push: one thread per rank waits (bounded) until that rank has published `epoch`.  Work
 * enqueued after this kernel on the same stream sees the complete table. */
__global__ void fleet_wait_kernel(const unsigned long long* flags, int world, unsigned long long epoch, int* err) {
  const int r = threadIdx.x;
  if (r < world) {
    const volatile unsigned long long* f = flags + r;
    long long spins = 0;
    while (*f < epoch) {
      __nanosleep(200);
      if (++spins > 5000000ll) { /* about a second: a peer died or never launched its update */
        *err = 1;
        break;
      }
    }
  }
  __threadfence_system();
}

}  // namespace b200nav
#endif

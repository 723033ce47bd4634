/*
 * b200nav_api.cu -- implementation of the C ABI declared in include/b200nav.h.
 *
 * Host logic only: contexts, device buffers, launch configuration, error plumbing.  All compute runs in the
 * sm_100a kernels of himm_kernels.cuh / vfh_kernels.cuh / grid_kernels.cuh; there is no CPU fallback.
 */
#include <cuda.h>
#include <dlfcn.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/b200nav.h"
#include "geometry.h"
#include "grid_kernels.cuh"
#include "himm_kernels.cuh"
#include "vfh_kernels.cuh"
#include "vfh_tables.h"

using namespace b200nav;

/* ================================================================================================================
 * Objects
 * ============================================================================================================== */

/* Event-pair pool for per-kernel device timing (b200nav_ctx_profile_*). */
enum { PROF_HIMM_PREP = 0, PROF_HIMM_TILE = 1, PROF_VFH = 2, PROF_HIMM_TILE_MW = 3, PROF_KINDS = 4 };
static const char* const kProfNames[PROF_KINDS] = {"himm_prep", "himm_tile", "vfh_update", "himm_tile_mw"};
struct ProfSlot {
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pairs; /* recorded, not yet read */
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> free_pairs;
  double total_ms = 0;
  int64_t count = 0;
};

struct b200nav_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool owns_stream = false;
  int64_t launches = 0;
  int sm_count = 0;
  bool profiling = false;
  unsigned prof_mask = ~0u;                /* kernel kinds that are timed while profiling (b200nav_ctx_profile_select) */
  cudaStream_t copy_stream = nullptr;      /* host->device copies pipelined with the prep kernel */
  std::vector<cudaEvent_t> copy_events;
  bool prep_done_valid = false;            /* copy_events[6] = "the last binning kernel has consumed the staged cloud" */
  cudaEvent_t fences[8] = {nullptr};       /* b200nav_ctx_fence / b200nav_ctx_wait */
  cudaEvent_t fences_side[8] = {nullptr};  /* the same tickets on the side stream (asynchronous VFH+ updates) */
  bool fence_has_side[8] = {false};
  int fence_seq = 0;
  /* asynchronous batched VFH+ updates run on a side stream so that the NEXT cycle's copies, L2 traffic and binning
   * kernel overlap them; `side_pending` = work on the side stream the main stream has not waited for yet */
  void* flush_buf = nullptr;               /* b200nav_ctx_flush_l2 scratch: [write half | read half] */
  size_t flush_cap = 0;
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_to_side = nullptr, ev_side_done = nullptr;
  cudaEvent_t ev_side_kernel = nullptr;    /* the side stream's last VFH+ KERNEL is done (its copy-out may still run) */
  bool side_pending = false;
  ProfSlot prof[PROF_KINDS];
  char err[512] = {0};
};

namespace {

thread_local char g_err[512] = {0};

int set_err(b200nav_ctx* ctx, int code, const char* fmt, ...) {
  char* dst = ctx ? ctx->err : g_err;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(dst, 512, fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(ctx, expr)                                                                              \
  do {                                                                                                   \
    cudaError_t e__ = (expr);                                                                            \
    if (e__ != cudaSuccess)                                                                              \
      return set_err(ctx, e__ == cudaErrorMemoryAllocation ? B200NAV_ENOMEM : B200NAV_ECUDA, "%s: %s", #expr, \
                     cudaGetErrorString(e__));                                                           \
  } while (0)

/* Grow-only device buffer. */
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = std::max(bytes, (size_t)4096);
    want = (want * 5) / 4;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

/* One layer of all robots; aliases (b200nav_grid_alias_layer) share the object.  Device format: cells.cuh. */
struct Layer {
  void* dev = nullptr;
  bool coded = true;
  /* [n_robots * n_tiles] free-column summaries of the HIMM tile kernel; all-zero = nothing known (CODED layers only
   * ever hold 0 or all-ones: the whole tile is free).  Every writer of the layer other than the tile kernel resets
   * it. */
  unsigned long long* free_cols = nullptr;
  float* fdev() const { return static_cast<float*>(dev); }
  uint8_t* cdev() const { return static_cast<uint8_t*>(dev); }
  ~Layer() {
    if (dev) cudaFree(dev);
    if (free_cols) cudaFree(free_cols);
  }
};

}  // namespace

struct b200nav_grid {
  b200nav_ctx* ctx = nullptr;
  GridDims dims;
  int n_robots = 1;
  std::vector<RobotGeom> geom_host;
  RobotGeom* geom_dev = nullptr;
  std::map<std::string, std::shared_ptr<Layer>> layers;
  DevBuf samples, segs, offsets, occ, stats, beam_masks, errflag, origins, clearbuf, touched, worklist, counters;
  /* cloud updates: offsets / origins double-buffered so that an asynchronous cycle can copy its own while the
   * previous cycle's tile kernel still reads the other slot */
  DevBuf offs2[2], orig2[2];
  cudaEvent_t tile_done[2] = {nullptr, nullptr};
  bool tile_done_set[2] = {false, false};
  int cloud_slot = 0;
  /* scan form: cached selection / regular offsets for the last b200nav_scan_info, staging of poses and ranges */
  DevBuf scan_sel, scan_offsets, scan_poses, scan_ranges;
  b200nav_scan_info scan_info_cached = b200nav_scan_info();
  bool scan_cache_valid = false;
  int scan_n_used = 0;
  float scan_increment_used = 0.f;
  DevBuf stage, convflag; /* float staging for upload / download of CODED layers; conversion "bad value" flag */
  int last_total = 0;
  size_t layer_elems() const { return (size_t)n_robots * dims.rows * dims.cols; }
};

/* Batched multi-GPU mode: the once-per-cycle all-gather of the robots' 16-byte commands (one process per GPU, robots
 * block-partitioned; no collective touches grids or scans).  NCCL is bound at run time (dlopen) so that the library
 * has no link-time dependency on it; the collective runs on its own stream and overlaps the next cycle's kernels. */
struct b200nav_fleet {
  b200nav_ctx* ctx = nullptr;
  void* lib = nullptr;
  struct NcclId { char internal[128]; };
  int (*InitRank)(void**, int, NcclId, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*ErrorString)(int) = nullptr;
  void* comm = nullptr;
  int rank = 0, world = 1;
  cudaStream_t stream = nullptr;
  cudaEvent_t ready[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr};
  bool pending[2] = {false, false};
  /* peer push (fused exchange): one IPC-shared region per rank = [2 slots][n_total commands] + [2 slots][world] flags
   * + [2 slots][world] consumer acknowledgements + a block counter + an error word */
  bool push = false;
  int n_local = 0, n_total = 0, row0 = 0;
  uint8_t* region = nullptr;
  size_t region_bytes = 0;
  uint8_t* peer_region[B200NAV_MAX_PEERS] = {nullptr};
  DevBuf cyc_local[2], cyc_table[2]; /* b200nav_fleet_cycle_async: this rank's rows / the gathered table, per slot */
  bool cyc_pending[2] = {false, false};
  unsigned long long epoch[2] = {0, 0};
  unsigned long long released[2] = {0, 0}; /* last epoch of the slot this rank has acknowledged as consumed */
  bool push_pending[2] = {false, false};
  b200nav_command* table(int p, int slot) const {
    return reinterpret_cast<b200nav_command*>(peer_region[p]) + (size_t)slot * n_total;
  }
  unsigned long long* flags(int p, int slot) const {
    return reinterpret_cast<unsigned long long*>(peer_region[p] + sizeof(b200nav_command) * 2 * (size_t)n_total) +
           (size_t)slot * world;
  }
  unsigned long long* acks(int p, int slot) const {
    return reinterpret_cast<unsigned long long*>(peer_region[p] + sizeof(b200nav_command) * 2 * (size_t)n_total) +
           (size_t)(2 + slot) * world;
  }
  unsigned int* counter() const {
    return reinterpret_cast<unsigned int*>(region + sizeof(b200nav_command) * 2 * (size_t)n_total +
                                           sizeof(unsigned long long) * 4 * (size_t)world);
  }
  int* errword() const { return reinterpret_cast<int*>(counter() + 1); }
};

struct b200nav_vfh {
  b200nav_ctx* ctx = nullptr;
  b200nav_vfh_params params;
  VfhTables tab;
  int n_robots = 1;
  VfhDev dev;
  /* device tables */
  float *d_dir = nullptr, *d_dist = nullptr, *d_base = nullptr;
  double* d_thr = nullptr;
  int16_t* d_kidx = nullptr;
  uint32_t* d_masks = nullptr;
  int32_t* d_mtr = nullptr;
  DevBuf in_buf, out_buf, ranges_buf;
  /* asynchronous batched updates: inputs double-buffered and copied on the context's copy stream */
  DevBuf in2[2];
  cudaEvent_t in_done[2] = {nullptr, nullptr}, in_ready = nullptr;
  bool in_done_set[2] = {false, false};
  int in_slot = 0;
  /* TMA descriptor cache: one per (layer pointer, geometry) */
  const float* tmap_layer = nullptr;
  int tmap_rows = 0, tmap_cols = 0, tmap_robots = 0, tmap_box_r = 0, tmap_box_c = 0;
  CUtensorMap tmap;
  bool tmap_valid = false;
  bool tma_disabled = false;
};

namespace {

/* Host-side wait for the main stream and the side stream. */
cudaError_t sync_raw(b200nav_ctx* ctx) {
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess && ctx->side_stream) e = cudaStreamSynchronize(ctx->side_stream);
  if (e == cudaSuccess) ctx->side_pending = false;
  return e;
}

/* Make the main stream wait for whatever is in flight on the side stream (see b200nav_ctx::side_stream).  Called by
 * everything that writes a layer, touches VFH+ state or frees buffers. */
int join_side(b200nav_ctx* ctx) {
  if (ctx->side_pending) {
    CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_side_done, 0));
    ctx->side_pending = false;
  }
  return B200NAV_OK;
}

/* The tile kernel only has to wait for the side stream's VFH+ kernel (the reader of the layer it rewrites), not for
 * the copy of the commands to the host that follows it there; side_pending stays set for everybody else. */
int join_side_layer_writer(b200nav_ctx* ctx) {
  if (ctx->side_pending) CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_side_kernel ? ctx->ev_side_kernel : ctx->ev_side_done, 0));
  return B200NAV_OK;
}

int sync_stream(b200nav_ctx* ctx) {
  CUDA_TRY(ctx, sync_raw(ctx));
  return B200NAV_OK;
}

int check_launch(b200nav_ctx* ctx, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err(ctx, B200NAV_ECUDA, "%s launch failed: %s", what, cudaGetErrorString(e));
  ctx->launches++;
  return B200NAV_OK;
}

/* RAII-ish helper: record an event pair around a launch when profiling is on. */
struct ProfScope {
  b200nav_ctx* ctx;
  int kind;
  std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};
  cudaStream_t st;
  ProfScope(b200nav_ctx* c, int k, cudaStream_t stream = nullptr) : ctx(c), kind(k), st(stream ? stream : c->stream) {
    if (!ctx->profiling || !((ctx->prof_mask >> kind) & 1u)) return;
    ProfSlot& s = ctx->prof[kind];
    if (!s.free_pairs.empty()) {
      ev = s.free_pairs.back();
      s.free_pairs.pop_back();
    } else if (cudaEventCreate(&ev.first) != cudaSuccess || cudaEventCreate(&ev.second) != cudaSuccess) {
      ev = {nullptr, nullptr};
      return;
    }
    cudaEventRecord(ev.first, st);
  }
  ~ProfScope() {
    if (!ev.first) return;
    cudaEventRecord(ev.second, st);
    ctx->prof[kind].pairs.push_back(ev);
  }
};

void prof_drain(b200nav_ctx* ctx) {
  for (int k = 0; k < PROF_KINDS; k++) {
    ProfSlot& s = ctx->prof[k];
    for (auto& p : s.pairs) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, p.first, p.second) == cudaSuccess) {
        s.total_ms += ms;
        s.count++;
      }
      s.free_pairs.push_back(p);
    }
    s.pairs.clear();
  }
}

Layer* find_layer(b200nav_grid* g, const char* name) {
  if (!name) return nullptr;
  auto it = g->layers.find(name);
  return it == g->layers.end() ? nullptr : it->second.get();
}

/* every distinct layer object once (aliases share theirs) */
std::vector<Layer*> unique_layers(b200nav_grid* g) {
  std::vector<Layer*> v;
  for (auto& kv : g->layers)
    if (std::find(v.begin(), v.end(), kv.second.get()) == v.end()) v.push_back(kv.second.get());
  return v;
}

size_t grid_tiles(const b200nav_grid* g) {
  return (size_t)((g->dims.rows + HIMM_TILE - 1) / HIMM_TILE) * ((g->dims.cols + HIMM_TILE - 1) / HIMM_TILE);
}

/* Forget what is known about free columns of `l` (robot < 0: all robots) - after any write that is not HIMM. */
int reset_free_cols(b200nav_grid* g, Layer* l, int robot) {
  if (!l->free_cols) return B200NAV_OK;
  const size_t nt = grid_tiles(g);
  if (robot < 0)
    CUDA_TRY(g->ctx, cudaMemsetAsync(l->free_cols, 0, sizeof(unsigned long long) * nt * g->n_robots, g->ctx->stream));
  else
    CUDA_TRY(g->ctx, cudaMemsetAsync(l->free_cols + nt * robot, 0, sizeof(unsigned long long) * nt, g->ctx->stream));
  return B200NAV_OK;
}

int fill_nan(b200nav_grid* g, float* p, size_t n) {
  const int threads = 256;
  const size_t blocks = std::min<size_t>((n + threads - 1) / threads, (size_t)g->ctx->sm_count * 16);
  grid_fill_kernel<<<(unsigned)std::max<size_t>(blocks, 1), threads, 0, g->ctx->stream>>>(p, n, nanf(""));
  return check_launch(g->ctx, "grid_fill_kernel");
}

size_t coded_robot_bytes(const b200nav_grid* g) { return grid_tiles(g) * (size_t)HIMM_TILE_BYTES; }
int grid_tiles_r(const b200nav_grid* g) { return (g->dims.rows + HIMM_TILE - 1) / HIMM_TILE; }

LayerRef layer_ref(const b200nav_grid* g, const Layer* l, int robot) {
  LayerRef r;
  r.coded = l->coded ? 1 : 0;
  r.rows = g->dims.rows;
  r.tiles_r = grid_tiles_r(g);
  r.base = l->coded ? static_cast<const void*>(l->cdev() + coded_robot_bytes(g) * robot)
                    : static_cast<const void*>(l->fdev() + (size_t)g->dims.rows * g->dims.cols * robot);
  return r;
}

unsigned conv_blocks(const b200nav_grid* g, size_t n, int threads) {
  return (unsigned)std::max<size_t>(1, std::min<size_t>((n + threads - 1) / threads, (size_t)g->ctx->sm_count * 16));
}

/* every cell of every robot := NaN */
int fill_layer_nan(b200nav_grid* g, Layer* l) {
  if (!l->coded) return fill_nan(g, l->fdev(), g->layer_elems());
  const size_t recs = grid_tiles(g) * (size_t)g->n_robots;
  coded_fill_kernel<<<conv_blocks(g, recs * HIMM_TILE_BYTES, 256), 256, 0, g->ctx->stream>>>(
      l->cdev(), recs, g->dims.rows, g->dims.cols, grid_tiles_r(g), (int)grid_tiles(g));
  return check_launch(g->ctx, "coded_fill_kernel");
}

/* CODED -> FLOAT, for good (foreign values arrived or the host wants the raw float pointer) */
int layer_to_float(b200nav_grid* g, Layer* l) {
  if (!l->coded) return B200NAV_OK;
  b200nav_ctx* ctx = g->ctx;
  float* f = nullptr;
  CUDA_TRY(ctx, cudaMalloc((void**)&f, g->layer_elems() * sizeof(float)));
  coded_to_float_kernel<<<conv_blocks(g, g->layer_elems(), 256), 256, 0, ctx->stream>>>(
      l->cdev(), f, g->dims.rows, g->dims.cols, grid_tiles_r(g), coded_robot_bytes(g), g->layer_elems());
  int rc = check_launch(ctx, "coded_to_float_kernel");
  if (rc == B200NAV_OK && sync_raw(ctx) != cudaSuccess) rc = B200NAV_ECUDA;
  if (rc) {
    cudaFree(f);
    return rc;
  }
  cudaFree(l->dev);
  l->dev = f;
  l->coded = false;
  return reset_free_cols(g, l, -1);
}

int alloc_layer(b200nav_grid* g, Layer* l, bool coded) {
  l->coded = coded;
  const size_t bytes = coded ? coded_robot_bytes(g) * g->n_robots : g->layer_elems() * sizeof(float);
  CUDA_TRY(g->ctx, cudaMalloc(&l->dev, bytes));
  const size_t fc_bytes = sizeof(unsigned long long) * grid_tiles(g) * g->n_robots;
  CUDA_TRY(g->ctx, cudaMalloc((void**)&l->free_cols, fc_bytes));
  CUDA_TRY(g->ctx, cudaMemsetAsync(l->free_cols, 0, fc_bytes, g->ctx->stream));
  return B200NAV_OK;
}

int upload_geom(b200nav_grid* g) {
  CUDA_TRY(g->ctx, cudaMemcpyAsync(g->geom_dev, g->geom_host.data(), sizeof(RobotGeom) * g->n_robots,
                                   cudaMemcpyHostToDevice, g->ctx->stream));
  /* geom_host may be modified right after return: make the copy complete first (tiny, rare). */
  return sync_stream(g->ctx);
}

constexpr int kVfhMaxSmem = 200 * 1024;
constexpr int kVfhManyWaves = 4096; /* robots per launch from which the occupancy-first VFH+ build is used */

/* ---- HIMM launch ------------------------------------------------------------------------------------------- */
constexpr int kSub = HIMM_TILE, kListCap = HIMM_CHUNK;
using TileCfg = HimmTileCfg<kSub, kListCap>;

/* Binning scratch: grow-only, kept all-zero between updates (the tile kernel clears what it consumes). */
int himm_reserve_masks(b200nav_grid* g, size_t n_tiles_total, int mask_words, size_t n_robot_tiles) {
  b200nav_ctx* ctx = g->ctx;
  const size_t mb = n_tiles_total * (size_t)mask_words * sizeof(uint32_t);
  if (mb > g->beam_masks.cap) {
    CUDA_TRY(ctx, sync_raw(ctx));
    CUDA_TRY(ctx, g->beam_masks.reserve(mb));
    CUDA_TRY(ctx, cudaMemsetAsync(g->beam_masks.p, 0, g->beam_masks.cap, ctx->stream));
  }
  if (!g->errflag.p) {
    CUDA_TRY(ctx, g->errflag.reserve(sizeof(int)));
    CUDA_TRY(ctx, cudaMemsetAsync(g->errflag.p, 0, sizeof(int), ctx->stream));
    CUDA_TRY(ctx, g->counters.reserve(8 * sizeof(int)));
    CUDA_TRY(ctx, cudaMemsetAsync(g->counters.p, 0, g->counters.cap, ctx->stream));
  }
  if (n_robot_tiles * sizeof(uint32_t) > g->touched.cap) {
    CUDA_TRY(ctx, sync_raw(ctx));
    CUDA_TRY(ctx, g->touched.reserve(n_robot_tiles * sizeof(uint32_t)));
    CUDA_TRY(ctx, cudaMemsetAsync(g->touched.p, 0, g->touched.cap, ctx->stream));
    CUDA_TRY(ctx, g->worklist.reserve(n_robot_tiles * sizeof(int)));
  }
  return B200NAV_OK;
}

struct CloudIn {
  const double* origins = nullptr;
  const float* xy = nullptr;
  const uint8_t* clear_end = nullptr;
  /* scan form */
  const float* scan_ranges = nullptr;
  const double* scan_poses = nullptr;
  const int32_t* scan_sel = nullptr;
  ScanModel scan = ScanModel();
};

int himm_setup(b200nav_grid* g, Layer* lay, const b200nav_sample* dev_samples, const int32_t* dev_offsets,
               int robot0, int n_active, int single_n, int total, int max_per_robot, CloudIn cloud, HimmArgs& a) {
  b200nav_ctx* ctx = g->ctx;
  CUDA_TRY(ctx, g->segs.reserve(sizeof(BeamSeg) * (size_t)total));
  a.dims = g->dims;
  a.geom = g->geom_dev;
  a.layer = lay->dev;
  a.coded = lay->coded ? 1 : 0;
  a.free_cols = lay->free_cols + grid_tiles(g) * (size_t)robot0;
  a.samples = dev_samples;
  a.origins = cloud.origins;
  a.xy = reinterpret_cast<const float2*>(cloud.xy);
  a.clear_end = cloud.clear_end;
  a.scan_ranges = cloud.scan_ranges;
  a.scan_poses = cloud.scan_poses;
  a.scan_sel = cloud.scan_sel;
  a.scan = cloud.scan;
  a.offsets = dev_offsets;
  a.segs = static_cast<BeamSeg*>(g->segs.p);
  a.robot0 = robot0;
  a.n_active = n_active;
  a.single_n = single_n;
  a.total = total;
  a.beam_lo = 0;
  a.beam_hi = total;
  a.rel_lo = 0;
  a.rel_hi = n_active;
  a.max_per_robot = max_per_robot;
  a.tiles_r = (g->dims.rows + TileCfg::kTileR - 1) / TileCfg::kTileR;
  a.tiles_c = (g->dims.cols + TileCfg::kTileC - 1) / TileCfg::kTileC;
  a.chunk_beams = std::min(HIMM_CHUNK, std::max(32, (max_per_robot + 31) & ~31));
  a.mask_words = a.chunk_beams / 32;
  a.n_chunks = std::max(1, (max_per_robot + a.chunk_beams - 1) / a.chunk_beams);
  const size_t n_tiles_total = (size_t)n_active * a.n_chunks * a.tiles_r * a.tiles_c;
  const size_t n_robot_tiles = (size_t)n_active * a.tiles_r * a.tiles_c;
  if (n_tiles_total * (size_t)a.mask_words * sizeof(uint32_t) > ((size_t)8 << 30))
    return set_err(ctx, B200NAV_ERANGE,
                   "binning scratch would need %zu MiB (robots x samples-per-robot/2048 x tiles): split the update",
                   (n_tiles_total * (size_t)a.mask_words * sizeof(uint32_t)) >> 20);
  int rc = himm_reserve_masks(g, n_tiles_total, a.mask_words, n_robot_tiles);
  if (rc) return rc;
  a.beam_masks = static_cast<uint32_t*>(g->beam_masks.p);
  a.error_flag = static_cast<int*>(g->errflag.p);
  a.touched = static_cast<uint32_t*>(g->touched.p);
  a.worklist = static_cast<int*>(g->worklist.p);
  a.counters = static_cast<int*>(g->counters.p);
  a.worklist_cap = (int)n_robot_tiles;
  a.mw_all = 0;
  a.defer_first_touch = n_active < kVfhManyWaves ? 1 : 0; /* see himm_prep_kernel; same threshold as the VFH+ builds */
  g->last_total = total;
  return B200NAV_OK;
}

/* K0 for beams [beam_lo, beam_hi) of robots [rel_lo, rel_hi) */
int himm_launch_prep(b200nav_grid* g, HimmArgs a, int beam_lo, int beam_hi, int rel_lo, int rel_hi) {
  b200nav_ctx* ctx = g->ctx;
  if (beam_hi <= beam_lo) return B200NAV_OK;
  ctx->prep_done_valid = false; /* a new reader of the staging buffers: the asynchronous paths re-arm it after the launch */
  a.beam_lo = beam_lo;
  a.beam_hi = beam_hi;
  a.rel_lo = rel_lo;
  a.rel_hi = rel_hi;
  {
    ProfScope ps(ctx, PROF_HIMM_PREP);
    const int n_rel = rel_hi - rel_lo;
    const unsigned gy = (unsigned)std::min(n_rel, 32768), gz = (unsigned)((n_rel + 32767) / 32768);
    dim3 grid((unsigned)std::max(1, (a.max_per_robot + 127) / 128), gy, gz);
    himm_prep_kernel<<<grid, 128, 0, ctx->stream>>>(a);
  }
  return check_launch(ctx, "himm_prep_kernel");
}

#ifndef B200NAV_HEAVY_WARPS
#define B200NAV_HEAVY_WARPS 8 /* A/B on C2 / C3 (tile phase): 2 warps 47 / 137 us, 4: 33 / 82, 6: 32 / 69, 8: 30 / 61 */
#endif
constexpr int kHeavyWarps = B200NAV_HEAVY_WARPS; /* warps per CTA of the multi-warp tile kernel */
constexpr int kMwAllRobots = 32; /* fleets up to this size: every tile item on a multi-warp CTA */

int himm_launch_tile(b200nav_grid* g, const HimmArgs& a_in) {
  b200nav_ctx* ctx = g->ctx;
  int jrc = join_side_layer_writer(ctx); /* a VFH+ update on the side stream may still read the layer this kernel rewrites */
  if (jrc) return jrc;
  HimmArgs a = a_in;
  const size_t smem = TileCfg::kTileBytes + sizeof(uint16_t) * (size_t)a.chunk_beams;
  /* The tile that holds a scan's own origin sees every beam of the scan: walked by one warp it is the critical path
   * of a small update (a single robot: 0.057 of 0.095 ms per scan).  A fleet of up to kMwAllRobots robots cannot fill
   * the GPU with one warp per tile anyway: there every touched tile gets a CTA of kHeavyWarps warps that share the
   * tile as a wavefront pipeline over the beam batches (himm_tile_coded_mw_kernel / himm_apply_list_pipe: no
   * replicated work) and the one-warp kernel is not launched.  Large fleets are throughput bound and keep one warp
   * per tile (measured: multi-warp CTAs for the heavy items of 1024 robots cost 0.33 instead of 0.28 ms per cycle).
   * B200NAV_MW_HEAVY=0: one warp per tile always; =2: multi-warp CTAs always (A/B aid). */
  static const char* mw_env = getenv("B200NAV_MW_HEAVY");
  const int mw_mode = mw_env ? atoi(mw_env) : 1;
  a.mw_all = (a.coded && (mw_mode == 2 || (mw_mode == 1 && a.n_active <= kMwAllRobots))) ? 1 : 0;
  if (a.mw_all) {
    ProfScope ps(ctx, PROF_HIMM_TILE_MW);
    const unsigned blocks = (unsigned)std::max<size_t>(1, std::min<size_t>((size_t)a.worklist_cap, (size_t)ctx->sm_count * (32 / kHeavyWarps)));
    himm_tile_coded_mw_kernel<kListCap, kHeavyWarps><<<blocks, 32 * kHeavyWarps, smem, ctx->stream>>>(a);
    return check_launch(ctx, "himm_tile_coded_mw_kernel");
  }
  /* persistent one-warp CTAs, never more than there are tiles */
  /* 30 one-warp CTAs fit an SM (shared memory); 28 resident ones measured best on C4 (tile walk 0.170 ms against
   * 0.176 with 30 and 0.177 with a grid of 32 per SM, 0.184 with 20): every CTA is resident from the start and the
   * warps that own the heavy origin tiles get a larger share of the issue slots.  B200NAV_TILE_CTAS_PER_SM: A/B aid. */
  static const char* cta_env = getenv("B200NAV_TILE_CTAS_PER_SM");
  const int per_sm = cta_env && atoi(cta_env) > 0 ? atoi(cta_env) : 28;
  dim3 grid((unsigned)std::min<size_t>((size_t)a.worklist_cap, (size_t)ctx->sm_count * per_sm));
  {
    ProfScope ps(ctx, PROF_HIMM_TILE);
    if (a.coded) himm_tile_coded_kernel<kListCap><<<grid, TileCfg::kThreads, smem, ctx->stream>>>(a);
    else himm_tile_kernel<kSub, kListCap><<<grid, TileCfg::kThreads, smem, ctx->stream>>>(a);
  }
  return check_launch(ctx, a.coded ? "himm_tile_coded_kernel" : "himm_tile_kernel");
}

int himm_launch(b200nav_grid* g, Layer* lay, const b200nav_sample* dev_samples, const int32_t* dev_offsets,
                int robot0, int n_active, int single_n, int total, int max_per_robot, CloudIn cloud = CloudIn()) {
  if (total <= 0) return B200NAV_OK;
  HimmArgs a;
  int rc = himm_setup(g, lay, dev_samples, dev_offsets, robot0, n_active, single_n, total, max_per_robot, cloud, a);
  if (rc) return rc;
  rc = himm_launch_prep(g, a, 0, total, 0, n_active);
  if (rc) return rc;
  return himm_launch_tile(g, a);
}

/* Reports (and clears) the device-side "more samples than declared" flag; call after a stream sync. */
int himm_check_error_flag(b200nav_grid* g) {
  if (!g->errflag.p) return B200NAV_OK;
  int flag = 0;
  CUDA_TRY(g->ctx, cudaMemcpy(&flag, g->errflag.p, sizeof(int), cudaMemcpyDeviceToHost));
  if (flag == 2) {
    cudaMemset(g->errflag.p, 0, sizeof(int));
    return set_err(g->ctx, B200NAV_ECUDA,
                   "the multi-warp tile kernel gave up waiting for a predecessor batch (internal error): the layer of "
                   "this update is not valid");
  }
  if (flag) {
    cudaMemset(g->errflag.p, 0, sizeof(int));
    return set_err(g->ctx, B200NAV_ERANGE,
                   "a robot had more samples than max_samples_per_robot: its excess samples were ignored and the "
                   "binning scratch may hold stale bits (destroy the grid)");
  }
  return B200NAV_OK;
}

void host_touch(const b200nav_sample* s, int n, double* bbox) {
  /* MapUpdater::touch (map_updater.h:73-78) over start and end of every sample */
  for (int i = 0; i < n; i++) {
    bbox[0] = std::min(bbox[0], s[i].sx);
    bbox[1] = std::min(bbox[1], s[i].sy);
    bbox[2] = std::max(bbox[2], s[i].sx);
    bbox[3] = std::max(bbox[3], s[i].sy);
    bbox[0] = std::min(bbox[0], s[i].ex);
    bbox[1] = std::min(bbox[1], s[i].ey);
    bbox[2] = std::max(bbox[2], s[i].ex);
    bbox[3] = std::max(bbox[3], s[i].ey);
  }
}

/* ---- VFH helpers -------------------------------------------------------------------------------------------- */
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_tmapEncodeTiled get_encode_fn() {
  static PFN_tmapEncodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
    cudaGetLastError();
  }
  return fn;
}

/* Window box for the VFH stage-R TMA load: (box_r x box_c) floats starting at the window's top-left cell. */
void vfh_window_box(const b200nav_vfh* v, const b200nav_grid* g, int& box_r, int& box_c) {
  const int n = (int)ceil(v->tab.c.submap_length / g->dims.res) + 2;
  box_c = std::min(n, g->dims.cols);
  box_r = std::min((n + 3 + 3) & ~3, (g->dims.rows + 3) & ~3); /* +3: the box starts at a 4-float aligned row */
}

bool vfh_prepare_tmap(b200nav_vfh* v, b200nav_grid* g, const float* layer, int box_r, int box_c) {
  static const bool env_off = getenv("B200NAV_DISABLE_TMA") != nullptr; /* debugging aid */
  if (v->tma_disabled || env_off) return false;
  if (g->dims.rows % 4 != 0 || box_r > 256 || box_c > 256) return false;
  if (v->tmap_valid && v->tmap_layer == layer && v->tmap_rows == g->dims.rows && v->tmap_cols == g->dims.cols &&
      v->tmap_robots == g->n_robots && v->tmap_box_r == box_r && v->tmap_box_c == box_c)
    return true;
  PFN_tmapEncodeTiled enc = get_encode_fn();
  if (!enc) return false;
  const cuuint64_t gdim[3] = {(cuuint64_t)g->dims.rows, (cuuint64_t)g->dims.cols, (cuuint64_t)g->n_robots};
  const cuuint64_t gstride[2] = {(cuuint64_t)g->dims.rows * sizeof(float),
                                 (cuuint64_t)g->dims.rows * g->dims.cols * sizeof(float)};
  const cuuint32_t box[3] = {(cuuint32_t)box_r, (cuuint32_t)box_c, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = enc(&v->tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(layer), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA);
  if (r != CUDA_SUCCESS) return false;
  v->tmap_valid = true;
  v->tmap_layer = layer;
  v->tmap_rows = g->dims.rows;
  v->tmap_cols = g->dims.cols;
  v->tmap_robots = g->n_robots;
  v->tmap_box_r = box_r;
  v->tmap_box_c = box_c;
  return true;
}

size_t vfh_smem_bytes(const b200nav_vfh* v, bool from_grid, int box_r, int box_c) {
  size_t b = from_grid ? (((size_t)box_r * box_c * sizeof(float) + 127) & ~(size_t)127) : 0;
  b += sizeof(double) * (B200NAV_NRANGES + 1);
  b += sizeof(float) * ((v->tab.c.hist_size + 1) & ~1);
  b += sizeof(uint16_t) * (size_t)v->dev.nf + 16;
  return b;
}

int vfh_launch(b200nav_vfh* v, b200nav_grid* g, const Layer* layer, const b200nav_vfh_input* dev_in,
               const double* dev_ranges, b200nav_command* dev_out, int robot0, int n, cudaStream_t stream = nullptr,
               const VfhPush* push_in = nullptr) {
  b200nav_ctx* ctx = v->ctx;
  if (!stream) stream = ctx->stream;
  VfhPush push;
  memset(&push, 0, sizeof(push));
  if (push_in) push = *push_in;
  VfhGridArgs ga;
  memset(&ga, 0, sizeof(ga));
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  size_t smem;
  if (g) {
    int box_r, box_c;
    vfh_window_box(v, g, box_r, box_c);
    ga.dims = g->dims;
    ga.geom = g->geom_dev;
    ga.layer = layer->dev;
    ga.coded = layer->coded ? 1 : 0;
    ga.tiles_r = grid_tiles_r(g);
    ga.tiles_c = (g->dims.cols + HIMM_TILE - 1) / HIMM_TILE;
    ga.box_r = box_r;
    ga.box_c = box_c;
    /* the tiled TMA window load works on the FLOAT layout; CODED layers are read through cells.cuh */
    ga.use_tma = (!layer->coded && vfh_prepare_tmap(v, g, layer->fdev(), box_r, box_c)) ? 1 : 0;
    if (ga.use_tma) tm = v->tmap;
    smem = vfh_smem_bytes(v, true, box_r, box_c);
    if (smem > (size_t)kVfhMaxSmem) return set_err(ctx, B200NAV_ERANGE, "VFH window too large for shared memory (%zu B)", smem);
    auto kern = n >= kVfhManyWaves ? vfh_update_kernel<true, 16> : vfh_update_kernel<true, 1>;
    ProfScope ps(ctx, PROF_VFH, stream);
    kern<<<n, B200NAV_VFH_THREADS, smem, stream>>>(v->dev, ga, tm, dev_in, nullptr, dev_out, robot0, push);
  } else {
    smem = vfh_smem_bytes(v, false, 0, 0);
    auto kern = n >= kVfhManyWaves ? vfh_update_kernel<false, 16> : vfh_update_kernel<false, 1>;
    ProfScope ps(ctx, PROF_VFH, stream);
    kern<<<n, B200NAV_VFH_THREADS, smem, stream>>>(v->dev, ga, tm, dev_in, dev_ranges, dev_out, robot0, push);
  }
  return check_launch(ctx, "vfh_update_kernel");
}

/* Opt in to > 48 KB dynamic shared memory once per device (attributes are per device). */
int configure_kernels(b200nav_ctx* ctx) {
  CUDA_TRY(ctx, cudaFuncSetAttribute(himm_tile_kernel<kSub, kListCap>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TileCfg::kSmemBytes));
  CUDA_TRY(ctx, cudaFuncSetAttribute(himm_tile_coded_kernel<kListCap>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TileCfg::kSmemBytes));
  CUDA_TRY(ctx, cudaFuncSetAttribute(himm_tile_coded_mw_kernel<kListCap, kHeavyWarps>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TileCfg::kSmemBytes));
  CUDA_TRY(ctx, cudaFuncSetAttribute(vfh_update_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kVfhMaxSmem));
  CUDA_TRY(ctx, cudaFuncSetAttribute(vfh_update_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kVfhMaxSmem));
  CUDA_TRY(ctx, cudaFuncSetAttribute(vfh_update_kernel<true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kVfhMaxSmem));
  CUDA_TRY(ctx, cudaFuncSetAttribute(vfh_update_kernel<false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kVfhMaxSmem));
  return B200NAV_OK;
}

template <class T>
int upload_vec(b200nav_ctx* ctx, T** dst, const std::vector<T>& src) {
  if (*dst) cudaFree(*dst);
  *dst = nullptr;
  CUDA_TRY(ctx, cudaMalloc((void**)dst, std::max<size_t>(src.size(), 1) * sizeof(T)));
  if (!src.empty()) CUDA_TRY(ctx, cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
  return B200NAV_OK;
}

}  // namespace

/* ================================================================================================================
 * Context
 * ============================================================================================================== */
extern "C" {

int b200nav_ctx_create(int device, void* cuda_stream, b200nav_ctx** out) {
  if (!out) return set_err(nullptr, B200NAV_EINVAL, "out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return set_err(nullptr, B200NAV_ENODEVICE, "no CUDA device (%s); this library has no CPU fallback",
                   e != cudaSuccess ? cudaGetErrorString(e) : "count == 0");
  if (device < 0 || device >= count) return set_err(nullptr, B200NAV_EINVAL, "device %d out of range [0,%d)", device, count);
  std::unique_ptr<b200nav_ctx> ctx(new b200nav_ctx());
  ctx->device = device;
  CUDA_TRY(nullptr, cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(nullptr, cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return set_err(nullptr, B200NAV_ENODEVICE, "device %d is sm_%d%d; kernels are built for sm_100a only", device,
                   prop.major, prop.minor);
  ctx->sm_count = prop.multiProcessorCount;
  if (cuda_stream) {
    ctx->stream = static_cast<cudaStream_t>(cuda_stream);
  } else {
    CUDA_TRY(nullptr, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->owns_stream = true;
  }
  int rc = configure_kernels(ctx.get());
  if (rc) {
    snprintf(g_err, sizeof(g_err), "%s", ctx->err);
    if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
    return rc;
  }
  *out = ctx.release();
  return B200NAV_OK;
}

int b200nav_ctx_destroy(b200nav_ctx* ctx) {
  if (!ctx) return B200NAV_OK;
  cudaSetDevice(ctx->device);
  sync_raw(ctx);
  if (ctx->side_stream) cudaStreamSynchronize(ctx->side_stream);
  prof_drain(ctx);
  for (int k = 0; k < PROF_KINDS; k++)
    for (auto& p : ctx->prof[k].free_pairs) {
      cudaEventDestroy(p.first);
      cudaEventDestroy(p.second);
    }
  for (cudaEvent_t e : ctx->copy_events) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->fences)
    if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->fences_side)
    if (e) cudaEventDestroy(e);
  if (ctx->side_stream) {
    cudaStreamSynchronize(ctx->side_stream);
    cudaStreamDestroy(ctx->side_stream);
  }
  if (ctx->flush_buf) cudaFree(ctx->flush_buf);
  if (ctx->ev_to_side) cudaEventDestroy(ctx->ev_to_side);
  if (ctx->ev_side_done) cudaEventDestroy(ctx->ev_side_done);
  if (ctx->ev_side_kernel) cudaEventDestroy(ctx->ev_side_kernel);
  if (ctx->copy_stream) {
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamDestroy(ctx->copy_stream);
  }
  if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return B200NAV_OK;
}

int b200nav_ctx_synchronize(b200nav_ctx* ctx) {
  if (!ctx) return B200NAV_EINVAL;
  return sync_stream(ctx);
}

int b200nav_ctx_flush_l2(b200nav_ctx* ctx, size_t write_bytes, size_t read_bytes) {
  if (!ctx) return B200NAV_EINVAL;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const size_t w16 = write_bytes / 16, r16 = read_bytes / 16;
  if ((w16 + r16) * 16 > ctx->flush_cap) {
    CUDA_TRY(ctx, sync_raw(ctx));
    if (ctx->flush_buf) cudaFree(ctx->flush_buf);
    ctx->flush_buf = nullptr;
    ctx->flush_cap = 0;
    CUDA_TRY(ctx, cudaMalloc(&ctx->flush_buf, (w16 + r16) * 16 + 16));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->flush_buf, 0, (w16 + r16) * 16 + 16, ctx->stream));
    ctx->flush_cap = (w16 + r16) * 16;
  }
  uint4* base = static_cast<uint4*>(ctx->flush_buf);
  unsigned* sink = reinterpret_cast<unsigned*>(base + w16 + r16);
  const unsigned blocks = (unsigned)ctx->sm_count * 8;
  if (w16) l2_flush_kernel<<<blocks, 256, 0, ctx->stream>>>(base, w16, 1, sink);
  if (r16) l2_flush_kernel<<<blocks, 256, 0, ctx->stream>>>(base + w16, r16, 2, sink);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err(ctx, B200NAV_ECUDA, "l2_flush_kernel: %s", cudaGetErrorString(e));
  return B200NAV_OK; /* not counted in launch_count: a measurement aid, not part of the path */
}

int b200nav_ctx_calibrate_red(b200nav_ctx* ctx, size_t buffer_bytes, double* reds_per_second) {
  if (!ctx || !reds_per_second || buffer_bytes < 4096) return B200NAV_EINVAL;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  size_t words = 1;
  while (words * 2 * sizeof(unsigned) <= buffer_bytes) words *= 2; /* a power of two: the index is a mask */
  if (words > (size_t)1 << 32) words = (size_t)1 << 32;
  unsigned* buf = nullptr;
  CUDA_TRY(ctx, cudaMalloc(&buf, words * sizeof(unsigned)));
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  int rc = B200NAV_OK;
  float ms = 0.0f;
  const unsigned blocks = (unsigned)ctx->sm_count * 8;
  const int iters = 256;
  auto step = [&](cudaError_t e, const char* what) {
    if (e != cudaSuccess && rc == B200NAV_OK) rc = set_err(ctx, B200NAV_ECUDA, "%s: %s", what, cudaGetErrorString(e));
  };
  step(cudaMemsetAsync(buf, 0, words * sizeof(unsigned), ctx->stream), "cudaMemsetAsync");
  step(cudaEventCreate(&e0), "cudaEventCreate");
  step(cudaEventCreate(&e1), "cudaEventCreate");
  if (rc == B200NAV_OK) {
    red_calibration_kernel<<<blocks, 256, 0, ctx->stream>>>(buf, (unsigned)(words - 1), iters, 1u); /* warm-up */
    step(cudaEventRecord(e0, ctx->stream), "cudaEventRecord");
    red_calibration_kernel<<<blocks, 256, 0, ctx->stream>>>(buf, (unsigned)(words - 1), iters, 2u);
    step(cudaEventRecord(e1, ctx->stream), "cudaEventRecord");
    step(cudaGetLastError(), "red_calibration_kernel");
    step(cudaEventSynchronize(e1), "cudaEventSynchronize");
    step(cudaEventElapsedTime(&ms, e0, e1), "cudaEventElapsedTime");
  }
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  cudaFree(buf);
  if (rc != B200NAV_OK) return rc;
  *reds_per_second = (double)blocks * 256.0 * (double)iters / ((double)ms * 1e-3);
  return B200NAV_OK; /* not counted in launch_count: a measurement aid, not part of the path */
}

int b200nav_ctx_fence(b200nav_ctx* ctx, int* ticket) {
  if (!ctx || !ticket) return B200NAV_EINVAL;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const int slot = ctx->fence_seq & 7;
  cudaEvent_t& e = ctx->fences[slot];
  if (!e) CUDA_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  CUDA_TRY(ctx, cudaEventRecord(e, ctx->stream));
  ctx->fence_has_side[slot] = false;
  if (ctx->side_pending) { /* the ticket also covers the side stream, without serialising the two streams */
    cudaEvent_t& es = ctx->fences_side[slot];
    if (!es) CUDA_TRY(ctx, cudaEventCreateWithFlags(&es, cudaEventDisableTiming));
    CUDA_TRY(ctx, cudaEventRecord(es, ctx->side_stream));
    ctx->fence_has_side[slot] = true;
  }
  *ticket = ctx->fence_seq++;
  return B200NAV_OK;
}

int b200nav_ctx_wait(b200nav_ctx* ctx, int ticket) {
  if (!ctx || ticket < 0 || ticket >= ctx->fence_seq) return B200NAV_EINVAL;
  /* a slot that was re-armed since marks a LATER point of the stream: waiting for it is still correct */
  CUDA_TRY(ctx, cudaEventSynchronize(ctx->fences[ticket & 7]));
  if (ctx->fence_has_side[ticket & 7]) CUDA_TRY(ctx, cudaEventSynchronize(ctx->fences_side[ticket & 7]));
  return B200NAV_OK;
}

void* b200nav_ctx_stream(b200nav_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

const char* b200nav_last_error(b200nav_ctx* ctx) { return ctx ? ctx->err : g_err; }

int64_t b200nav_ctx_launch_count(b200nav_ctx* ctx) { return ctx ? ctx->launches : 0; }

int b200nav_ctx_profile_enable(b200nav_ctx* ctx, int enable) {
  if (!ctx) return B200NAV_EINVAL;
  int src = sync_stream(ctx);
  if (src) return src;
  prof_drain(ctx);
  ctx->profiling = enable != 0;
  if (enable) /* event pairs are only recycled at drain points: have enough ready so that timed launches create none */
    for (int k = 0; k < PROF_KINDS; k++)
      while (ctx->prof[k].free_pairs.size() < 512) {
        std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};
        if (cudaEventCreate(&ev.first) != cudaSuccess || cudaEventCreate(&ev.second) != cudaSuccess) break;
        ctx->prof[k].free_pairs.push_back(ev);
      }
  if (enable)
    for (int k = 0; k < PROF_KINDS; k++) {
      ctx->prof[k].total_ms = 0;
      ctx->prof[k].count = 0;
    }
  return B200NAV_OK;
}

int b200nav_ctx_profile_select(b200nav_ctx* ctx, const char* names) {
  if (!ctx) return B200NAV_EINVAL;
  if (!names || !*names) {
    ctx->prof_mask = ~0u;
    return B200NAV_OK;
  }
  unsigned mask = 0u;
  const char* p = names;
  while (*p) {
    const char* e = strchr(p, ',');
    const size_t len = e ? (size_t)(e - p) : strlen(p);
    bool found = false;
    for (int k = 0; k < PROF_KINDS; k++)
      if (strlen(kProfNames[k]) == len && strncmp(p, kProfNames[k], len) == 0) {
        mask |= 1u << k;
        found = true;
      }
    if (!found) return set_err(ctx, B200NAV_EINVAL, "profile_select: unknown kernel name in '%s'", names);
    p += len + (e ? 1 : 0);
  }
  ctx->prof_mask = mask;
  return B200NAV_OK;
}

int b200nav_ctx_profile_read(b200nav_ctx* ctx, const char* name, double* total_ms, int64_t* launches) {
  if (!ctx || !name) return B200NAV_EINVAL;
  int src = sync_stream(ctx);
  if (src) return src;
  prof_drain(ctx);
  for (int k = 0; k < PROF_KINDS; k++)
    if (strcmp(name, kProfNames[k]) == 0) {
      if (total_ms) *total_ms = ctx->prof[k].total_ms;
      if (launches) *launches = ctx->prof[k].count;
      return B200NAV_OK;
    }
  return set_err(ctx, B200NAV_EINVAL, "unknown kernel name '%s'", name);
}

/* ================================================================================================================
 * Grid
 * ============================================================================================================== */

int b200nav_grid_create(b200nav_ctx* ctx, double len_x, double len_y, double res, double pos_x, double pos_y,
                        int n_robots, b200nav_grid** out) {
  if (!ctx || !out) return B200NAV_EINVAL;
  *out = nullptr;
  if (!(len_x > 0) || !(len_y > 0) || !(res > 0) || n_robots < 1)
    return set_err(ctx, B200NAV_EINVAL, "grid_create: length/resolution must be > 0 and n_robots >= 1");
  const double fr = round(len_x / res), fc = round(len_y / res);
  if (fr < 1 || fc < 1 || fr > 32767 || fc > 32767)
    return set_err(ctx, B200NAV_ERANGE, "grid_create: %g x %g cells outside [1, 32767]", fr, fc);
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  std::unique_ptr<b200nav_grid> g(new b200nav_grid());
  g->ctx = ctx;
  g->dims.rows = (int)fr;
  g->dims.cols = (int)fc;
  g->dims.res = res;
  g->dims.len_x = (double)g->dims.rows * res;
  g->dims.len_y = (double)g->dims.cols * res;
  g->dims.rres = 1.0 / res;
  g->n_robots = n_robots;
  RobotGeom rg;
  rg.pos_x = pos_x;
  rg.pos_y = pos_y;
  rg.start0 = rg.start1 = 0;
  g->geom_host.assign(n_robots, rg);
  CUDA_TRY(ctx, cudaMalloc((void**)&g->geom_dev, sizeof(RobotGeom) * n_robots));
  int rc = upload_geom(g.get());
  if (rc) {
    cudaFree(g->geom_dev);
    return rc;
  }
  *out = g.release();
  return B200NAV_OK;
}

int b200nav_grid_destroy(b200nav_grid* g) {
  if (!g) return B200NAV_OK;
  cudaSetDevice(g->ctx->device);
  sync_raw(g->ctx);
  g->layers.clear(); /* frees the device buffers */
  g->stage.release();
  g->convflag.release();
  g->scan_sel.release();
  g->scan_offsets.release();
  g->scan_poses.release();
  g->scan_ranges.release();
  for (int i = 0; i < 2; i++) {
    g->offs2[i].release();
    g->orig2[i].release();
    if (g->tile_done[i]) cudaEventDestroy(g->tile_done[i]);
  }
  if (g->geom_dev) cudaFree(g->geom_dev);
  g->samples.release();
  g->segs.release();
  g->offsets.release();
  g->occ.release();
  g->stats.release();
  g->beam_masks.release();
  g->errflag.release();
  g->origins.release();
  g->clearbuf.release();
  g->touched.release();
  g->worklist.release();
  g->counters.release();
  delete g;
  return B200NAV_OK;
}

int b200nav_grid_size(const b200nav_grid* g, int* rows, int* cols, int* n_robots) {
  if (!g) return B200NAV_EINVAL;
  if (rows) *rows = g->dims.rows;
  if (cols) *cols = g->dims.cols;
  if (n_robots) *n_robots = g->n_robots;
  return B200NAV_OK;
}

int b200nav_grid_add_layer(b200nav_grid* g, const char* name) {
  if (!g || !name || !*name) return B200NAV_EINVAL;
  if (find_layer(g, name)) return B200NAV_OK; /* map_.exists(typeName) (map_updater.h:12) */
  CUDA_TRY(g->ctx, cudaSetDevice(g->ctx->device));
  static const bool float_layers = getenv("B200NAV_FLOAT_LAYERS") != nullptr; /* testing aid: start in FLOAT format */
  auto l = std::make_shared<Layer>();
  int rc = alloc_layer(g, l.get(), !float_layers);
  if (rc) return rc;
  rc = fill_layer_nan(g, l.get());
  if (rc) return rc;
  g->layers[name] = l;
  return B200NAV_OK;
}

int b200nav_grid_alias_layer(b200nav_grid* g, const char* alias, const char* target) {
  if (!g || !alias || !target) return B200NAV_EINVAL;
  Layer* t = find_layer(g, target);
  if (!t) return set_err(g->ctx, B200NAV_ENOLAYER, "no layer '%s'", target);
  if (find_layer(g, alias) == t) return B200NAV_OK;
  sync_raw(g->ctx); /* a layer replaced by the alias may still be in use */
  g->layers[alias] = g->layers[target];
  return B200NAV_OK;
}

int b200nav_grid_copy_layer(b200nav_grid* g, const char* dst, const char* src) {
  if (!g) return B200NAV_EINVAL;
  Layer *d = find_layer(g, dst), *s = find_layer(g, src);
  if (!d || !s) return set_err(g->ctx, B200NAV_ENOLAYER, "no layer '%s'", !d ? dst : src);
  if (d == s) return B200NAV_OK;
  { int jrc = join_side(g->ctx); if (jrc) return jrc; }
  if (d->coded != s->coded) { /* the destination takes the source's format */
    CUDA_TRY(g->ctx, sync_raw(g->ctx));
    void* nd = nullptr;
    const size_t nbytes = s->coded ? coded_robot_bytes(g) * g->n_robots : g->layer_elems() * sizeof(float);
    CUDA_TRY(g->ctx, cudaMalloc(&nd, nbytes));
    cudaFree(d->dev);
    d->dev = nd;
    d->coded = s->coded;
  }
  const size_t bytes = s->coded ? coded_robot_bytes(g) * g->n_robots : g->layer_elems() * sizeof(float);
  CUDA_TRY(g->ctx, cudaMemcpyAsync(d->dev, s->dev, bytes, cudaMemcpyDeviceToDevice, g->ctx->stream));
  /* the copy carries the source's free-column knowledge along */
  CUDA_TRY(g->ctx, cudaMemcpyAsync(d->free_cols, s->free_cols, sizeof(unsigned long long) * grid_tiles(g) * g->n_robots,
                                   cudaMemcpyDeviceToDevice, g->ctx->stream));
  return B200NAV_OK;
}

int b200nav_grid_compose_master(b200nav_grid* g, const char* dst, const char* range_layer, const char* laser_layer) {
  if (!g) return B200NAV_EINVAL;
  Layer *d = find_layer(g, dst), *a = find_layer(g, range_layer), *b = find_layer(g, laser_layer);
  if (!d || !a || !b) return set_err(g->ctx, B200NAV_ENOLAYER, "no layer '%s'", !d ? dst : (!a ? range_layer : laser_layer));
  if (d == a || d == b) return set_err(g->ctx, B200NAV_EINVAL, "compose: the destination must not alias a source layer");
  b200nav_ctx* ctx = g->ctx;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  { int jrc = join_side(ctx); if (jrc) return jrc; }
  if (d->coded) { /* sums leave the HIMM value set: the destination becomes a float layer (contents are overwritten) */
    CUDA_TRY(ctx, sync_raw(ctx));
    void* nd = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&nd, g->layer_elems() * sizeof(float)));
    cudaFree(d->dev);
    d->dev = nd;
    d->coded = false;
  }
  int rc = reset_free_cols(g, d, -1);
  if (rc) return rc;
  const size_t per = (size_t)g->dims.rows * g->dims.cols;
  grid_compose_kernel<<<conv_blocks(g, g->layer_elems(), 256), 256, 0, ctx->stream>>>(
      layer_ref(g, a, 0), a->coded ? coded_robot_bytes(g) : per * sizeof(float), layer_ref(g, b, 0),
      b->coded ? coded_robot_bytes(g) : per * sizeof(float), d->fdev(), g->dims.rows, g->dims.cols, g->layer_elems());
  return check_launch(ctx, "grid_compose_kernel");
}

int b200nav_grid_clear(b200nav_grid* g, const char* layer) {
  if (!g) return B200NAV_EINVAL;
  { int jrc = join_side(g->ctx); if (jrc) return jrc; }
  if (layer) {
    Layer* l = find_layer(g, layer);
    if (!l) return set_err(g->ctx, B200NAV_ENOLAYER, "no layer '%s'", layer);
    int rc = reset_free_cols(g, l, -1);
    if (rc) return rc;
    return fill_layer_nan(g, l);
  }
  for (Layer* l : unique_layers(g)) {
    int rc = reset_free_cols(g, l, -1);
    if (rc) return rc;
    rc = fill_layer_nan(g, l);
    if (rc) return rc;
  }
  return B200NAV_OK;
}

int b200nav_grid_upload(b200nav_grid* g, int robot, const char* layer, const float* colmajor) {
  if (!g || !colmajor || robot < 0 || robot >= g->n_robots) return B200NAV_EINVAL;
  Layer* l = find_layer(g, layer);
  if (!l) return set_err(g->ctx, B200NAV_ENOLAYER, "no layer '%s'", layer ? layer : "(null)");
  const size_t n = (size_t)g->dims.rows * g->dims.cols;
  b200nav_ctx* ctx = g->ctx;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  { int jrc = join_side(ctx); if (jrc) return jrc; }
  int rrc = reset_free_cols(g, l, robot);
  if (rrc) return rrc;
  if (l->coded) {
    /* float staging -> check that every value has a code -> encode into the robot's records */
    CUDA_TRY(ctx, g->stage.reserve(n * sizeof(float)));
    CUDA_TRY(ctx, g->convflag.reserve(sizeof(int)));
    CUDA_TRY(ctx, cudaMemsetAsync(g->convflag.p, 0, sizeof(int), ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(g->stage.p, colmajor, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    uint8_t* dst = l->cdev() + coded_robot_bytes(g) * robot;
    for (int pass = 0; pass < 2; pass++) {
      coded_from_float_kernel<<<conv_blocks(g, n, 256), 256, 0, ctx->stream>>>(
          static_cast<const float*>(g->stage.p), dst, g->dims.rows, g->dims.cols, grid_tiles_r(g), coded_robot_bytes(g),
          n, pass == 0 ? 1 : 0, static_cast<int*>(g->convflag.p));
      int rc = check_launch(ctx, "coded_from_float_kernel");
      if (rc) return rc;
      if (pass == 0) {
        int bad = 0;
        CUDA_TRY(ctx, cudaMemcpyAsync(&bad, g->convflag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, sync_raw(ctx));
        if (bad) { /* values outside the HIMM set: the layer leaves the coded format */
          rc = layer_to_float(g, l);
          if (rc) return rc;
          CUDA_TRY(ctx, cudaMemcpyAsync(l->fdev() + n * robot, g->stage.p, n * sizeof(float), cudaMemcpyDeviceToDevice,
                                        ctx->stream));
          break;
        }
      }
    }
    return sync_stream(ctx);
  }
  CUDA_TRY(ctx, cudaMemcpyAsync(l->fdev() + n * robot, colmajor, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  return sync_stream(ctx);
}

int b200nav_grid_download(b200nav_grid* g, int robot, const char* layer, float* colmajor) {
  if (!g || !colmajor || robot < 0 || robot >= g->n_robots) return B200NAV_EINVAL;
  Layer* l = find_layer(g, layer);
  if (!l) return set_err(g->ctx, B200NAV_ENOLAYER, "no layer '%s'", layer ? layer : "(null)");
  const size_t n = (size_t)g->dims.rows * g->dims.cols;
  b200nav_ctx* ctx = g->ctx;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const float* src = l->coded ? nullptr : l->fdev() + n * robot;
  if (l->coded) {
    CUDA_TRY(ctx, g->stage.reserve(n * sizeof(float)));
    coded_to_float_kernel<<<conv_blocks(g, n, 256), 256, 0, ctx->stream>>>(
        l->cdev() + coded_robot_bytes(g) * robot, static_cast<float*>(g->stage.p), g->dims.rows, g->dims.cols,
        grid_tiles_r(g), coded_robot_bytes(g), n);
    int rc = check_launch(ctx, "coded_to_float_kernel");
    if (rc) return rc;
    src = static_cast<const float*>(g->stage.p);
  }
  CUDA_TRY(ctx, cudaMemcpyAsync(colmajor, src, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  return sync_stream(ctx);
}

int b200nav_grid_set_geometry(b200nav_grid* g, int robot, double pos_x, double pos_y, int start0, int start1) {
  if (!g || robot < 0 || robot >= g->n_robots) return B200NAV_EINVAL;
  if (start0 < 0 || start0 >= g->dims.rows || start1 < 0 || start1 >= g->dims.cols)
    return set_err(g->ctx, B200NAV_EINVAL, "start index (%d,%d) outside the buffer", start0, start1);
  /* an asynchronous VFH+ update on the side stream may still read geom_dev[robot] for the previous cycle */
  { int jrc = join_side(g->ctx); if (jrc) return jrc; }
  RobotGeom& rg = g->geom_host[robot];
  rg.pos_x = pos_x;
  rg.pos_y = pos_y;
  rg.start0 = start0;
  rg.start1 = start1;
  CUDA_TRY(g->ctx, cudaMemcpyAsync(g->geom_dev + robot, &rg, sizeof(RobotGeom), cudaMemcpyHostToDevice,
                                   g->ctx->stream));
  return sync_stream(g->ctx);
}

int b200nav_grid_get_geometry(const b200nav_grid* g, int robot, double* pos_x, double* pos_y, int* start0,
                              int* start1) {
  if (!g || robot < 0 || robot >= g->n_robots) return B200NAV_EINVAL;
  const RobotGeom& rg = g->geom_host[robot];
  if (pos_x) *pos_x = rg.pos_x;
  if (pos_y) *pos_y = rg.pos_y;
  if (start0) *start0 = rg.start0;
  if (start1) *start1 = rg.start1;
  return B200NAV_OK;
}

int b200nav_grid_move(b200nav_grid* g, int robot, double x, double y, int* moved) {
  if (!g || robot < 0 || robot >= g->n_robots) return B200NAV_EINVAL;
  { int jrc = join_side(g->ctx); if (jrc) return jrc; }
  /* GridMap::move (GridMap.cpp:346-412).  The index/position bookkeeping is a handful of scalar operations and
   * stays on the host (the GridMap object owns its geometry); the strips that fall out of the map are NaN-filled
   * on the device in every layer. */
  RobotGeom& rg = g->geom_host[robot];
  const double res = g->dims.res;
  const double t[2] = {(x - rg.pos_x) / res, (y - rg.pos_y) / res};
  int shift[2];
  for (int i = 0; i < 2; i++) shift[i] = -(int)(t[i] + 0.5 * (t[i] > 0 ? 1 : -1));
  const int size[2] = {g->dims.rows, g->dims.cols};
  int start[2] = {rg.start0, rg.start1};
  const size_t per_robot = (size_t)g->dims.rows * g->dims.cols;
  auto clear_strip = [&](int axis, int index, int n) -> int {
    if (n <= 0) return B200NAV_OK;
    for (Layer* l : unique_layers(g)) {
      int rrc = reset_free_cols(g, l, robot);
      if (rrc) return rrc;
      const int r0 = axis == 0 ? index : 0, nr = axis == 0 ? n : g->dims.rows;
      const int c0 = axis == 1 ? index : 0, nc = axis == 1 ? n : g->dims.cols;
      dim3 grid((unsigned)((nr + 127) / 128), (unsigned)std::min(nc, 65535));
      if (l->coded)
        coded_fill_rect_kernel<<<grid, 128, 0, g->ctx->stream>>>(l->cdev() + coded_robot_bytes(g) * robot,
                                                                 grid_tiles_r(g), r0, nr, c0, nc, HIMM_CODE_NAN);
      else
        grid_fill_rect_kernel<<<grid, 128, 0, g->ctx->stream>>>(l->fdev() + per_robot * robot, g->dims.rows, r0, nr,
                                                                c0, nc, nanf(""));
      int rc = check_launch(g->ctx, "grid_fill_rect_kernel");
      if (rc) return rc;
    }
    return B200NAV_OK;
  };
  for (int i = 0; i < 2; i++) {
    if (shift[i] == 0) continue;
    if (abs(shift[i]) >= size[i]) {
      int rc = clear_strip(0, 0, g->dims.rows);
      if (rc) return rc;
    } else {
      const int sign = (shift[i] > 0 ? 1 : -1);
      const int start_index = start[i] - (sign < 0 ? 1 : 0);
      const int end_index = start_index - sign + shift[i];
      const int n_cells = abs(shift[i]);
      int index = (sign > 0 ? start_index : end_index);
      wrap_index(index, size[i]);
      if (index + n_cells <= size[i]) {
        int rc = clear_strip(i, index, n_cells);
        if (rc) return rc;
      } else {
        const int first_n = size[i] - index;
        int rc = clear_strip(i, index, first_n);
        if (rc) return rc;
        rc = clear_strip(i, 0, n_cells - first_n);
        if (rc) return rc;
      }
    }
  }
  rg.start0 += shift[0];
  rg.start1 += shift[1];
  wrap_index(rg.start0, g->dims.rows);
  wrap_index(rg.start1, g->dims.cols);
  rg.pos_x += (double)(-shift[0]) * res;
  rg.pos_y += (double)(-shift[1]) * res;
  if (moved) *moved = (shift[0] != 0 || shift[1] != 0) ? 1 : 0;
  CUDA_TRY(g->ctx, cudaMemcpyAsync(g->geom_dev + robot, &rg, sizeof(RobotGeom), cudaMemcpyHostToDevice,
                                   g->ctx->stream));
  return sync_stream(g->ctx);
}

int b200nav_grid_to_occupancy(b200nav_grid* g, int robot, const char* layer, float data_min, float data_max,
                              int8_t* out_host) {
  if (!g || !out_host || robot < 0 || robot >= g->n_robots) return B200NAV_EINVAL;
  Layer* l = find_layer(g, layer);
  if (!l) return set_err(g->ctx, B200NAV_ENOLAYER, "no layer '%s'", layer ? layer : "(null)");
  const size_t n = (size_t)g->dims.rows * g->dims.cols;
  CUDA_TRY(g->ctx, g->occ.reserve(n));
  const RobotGeom& rg = g->geom_host[robot];
  const int threads = 256;
  grid_to_occupancy_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, g->ctx->stream>>>(
      layer_ref(g, l, robot), g->dims.rows, g->dims.cols, rg.start0, rg.start1, data_min, data_max,
      static_cast<int8_t*>(g->occ.p));
  int rc = check_launch(g->ctx, "grid_to_occupancy_kernel");
  if (rc) return rc;
  CUDA_TRY(g->ctx, cudaMemcpyAsync(out_host, g->occ.p, n, cudaMemcpyDeviceToHost, g->ctx->stream));
  return sync_stream(g->ctx);
}

int b200nav_grid_query_blocked(b200nav_grid* g, int robot, const char* layer, const double* host_xy, int n,
                               double radius, uint8_t* out_host) {
  if (!g || robot < 0 || robot >= g->n_robots || n < 0 || (n > 0 && (!host_xy || !out_host)) || !(radius >= 0))
    return B200NAV_EINVAL;
  Layer* l = find_layer(g, layer);
  if (!l) return set_err(g->ctx, B200NAV_ENOLAYER, "no layer '%s'", layer ? layer : "(null)");
  if (n == 0) return B200NAV_OK;
  b200nav_ctx* ctx = g->ctx;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  CUDA_TRY(ctx, g->samples.reserve(sizeof(double) * 2 * (size_t)n));
  CUDA_TRY(ctx, g->occ.reserve((size_t)n));
  CUDA_TRY(ctx, cudaMemcpyAsync(g->samples.p, host_xy, sizeof(double) * 2 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  grid_blocked_kernel<<<(unsigned)((n * 32 + 127) / 128), 128, 0, ctx->stream>>>(
      g->dims, g->geom_host[robot], layer_ref(g, l, robot), static_cast<const double*>(g->samples.p), n, radius,
      static_cast<uint8_t*>(g->occ.p));
  int rc = check_launch(ctx, "grid_blocked_kernel");
  if (rc) return rc;
  CUDA_TRY(ctx, cudaMemcpyAsync(out_host, g->occ.p, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  return sync_stream(ctx);
}

int b200nav_grid_layer_written(b200nav_grid* g, const char* layer, int robot) {
  if (!g || robot >= g->n_robots) return B200NAV_EINVAL;
  Layer* l = find_layer(g, layer);
  if (!l) return set_err(g->ctx, B200NAV_ENOLAYER, "no layer '%s'", layer ? layer : "(null)");
  return reset_free_cols(g, l, robot);
}

void* b200nav_grid_layer_devptr(b200nav_grid* g, const char* layer) {
  if (!g) return nullptr;
  Layer* l = find_layer(g, layer);
  if (!l) return nullptr;
  /* the raw pointer is the reference's float layout: a CODED layer is converted (and stays FLOAT) */
  if (l->coded && layer_to_float(g, l) != B200NAV_OK) return nullptr;
  return l->dev;
}

int b200nav_grid_has_layer(const b200nav_grid* g, const char* layer) {
  if (!g || !layer) return 0;
  return g->layers.find(layer) != g->layers.end() ? 1 : 0;
}

int b200nav_grid_layer_format(b200nav_grid* g, const char* layer) {
  if (!g) return B200NAV_EINVAL;
  Layer* l = find_layer(g, layer);
  if (!l) return set_err(g->ctx, B200NAV_ENOLAYER, "no layer '%s'", layer ? layer : "(null)");
  return l->coded ? B200NAV_LAYER_CODED : B200NAV_LAYER_FLOAT;
}

/* ================================================================================================================
 * Fleet (multi-GPU command exchange)
 * ============================================================================================================== */

static void* nccl_open() {
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  return h;
}

int b200nav_fleet_unique_id(uint8_t* id128) {
  if (!id128) return B200NAV_EINVAL;
  void* h = nccl_open();
  if (!h) return set_err(nullptr, B200NAV_ENODEVICE, "libnccl.so.2 not found: %s", dlerror());
  auto get = reinterpret_cast<int (*)(void*)>(dlsym(h, "ncclGetUniqueId"));
  if (!get) return set_err(nullptr, B200NAV_ENODEVICE, "ncclGetUniqueId missing");
  const int r = get(id128);
  return r == 0 ? B200NAV_OK : set_err(nullptr, B200NAV_ECUDA, "ncclGetUniqueId failed (%d)", r);
}

int b200nav_fleet_create(b200nav_ctx* ctx, const uint8_t* id128, int rank, int world, b200nav_fleet** out) {
  if (!ctx || !id128 || !out || world < 1 || rank < 0 || rank >= world) return B200NAV_EINVAL;
  *out = nullptr;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  std::unique_ptr<b200nav_fleet> f(new b200nav_fleet());
  f->ctx = ctx;
  f->rank = rank;
  f->world = world;
  f->lib = nccl_open();
  if (!f->lib) return set_err(ctx, B200NAV_ENODEVICE, "libnccl.so.2 not found: %s", dlerror());
  f->InitRank = reinterpret_cast<decltype(f->InitRank)>(dlsym(f->lib, "ncclCommInitRank"));
  f->AllGather = reinterpret_cast<decltype(f->AllGather)>(dlsym(f->lib, "ncclAllGather"));
  f->CommDestroy = reinterpret_cast<decltype(f->CommDestroy)>(dlsym(f->lib, "ncclCommDestroy"));
  f->ErrorString = reinterpret_cast<decltype(f->ErrorString)>(dlsym(f->lib, "ncclGetErrorString"));
  if (!f->InitRank || !f->AllGather || !f->CommDestroy) return set_err(ctx, B200NAV_ENODEVICE, "NCCL symbols missing");
  b200nav_fleet::NcclId id;
  memcpy(id.internal, id128, sizeof(id.internal));
  const int r = f->InitRank(&f->comm, world, id, rank);
  if (r != 0) return set_err(ctx, B200NAV_ECUDA, "ncclCommInitRank: %s", f->ErrorString ? f->ErrorString(r) : "?");
  CUDA_TRY(ctx, cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; i++) {
    CUDA_TRY(ctx, cudaEventCreateWithFlags(&f->ready[i], cudaEventDisableTiming));
    CUDA_TRY(ctx, cudaEventCreateWithFlags(&f->done[i], cudaEventDisableTiming));
  }
  *out = f.release();
  return B200NAV_OK;
}

int b200nav_fleet_gather_async(b200nav_fleet* f, int slot, const void* dev_local, void* dev_table,
                               size_t bytes_per_rank) {
  if (!f || slot < 0 || slot > 1 || !dev_local || !dev_table) return B200NAV_EINVAL;
  b200nav_ctx* ctx = f->ctx;
  { int jrc = join_side(ctx); if (jrc) return jrc; }
  /* after everything enqueued on the context's stream so far (the VFH+ kernel that wrote dev_local) */
  CUDA_TRY(ctx, cudaEventRecord(f->ready[slot], ctx->stream));
  CUDA_TRY(ctx, cudaStreamWaitEvent(f->stream, f->ready[slot], 0));
  const int r = f->AllGather(dev_local, dev_table, bytes_per_rank, /*ncclUint8*/ 1, f->comm, f->stream);
  if (r != 0) return set_err(ctx, B200NAV_ECUDA, "ncclAllGather: %s", f->ErrorString ? f->ErrorString(r) : "?");
  CUDA_TRY(ctx, cudaEventRecord(f->done[slot], f->stream));
  f->pending[slot] = true;
  return B200NAV_OK;
}

int b200nav_fleet_wait(b200nav_fleet* f, int slot) {
  if (!f || slot > 1) return B200NAV_EINVAL;
  for (int s = (slot < 0 ? 0 : slot); s <= (slot < 0 ? 1 : slot); s++)
    if (f->push_pending[s]) { /* peer push: wait (on the stream) until every rank has published this slot's epoch */
      fleet_wait_kernel<<<1, 32, 0, f->ctx->stream>>>(f->flags(f->rank, s), f->world, f->epoch[s], f->errword());
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return set_err(f->ctx, B200NAV_ECUDA, "fleet_wait_kernel: %s", cudaGetErrorString(e));
      f->push_pending[s] = false;
    }
  for (int s = (slot < 0 ? 0 : slot); s <= (slot < 0 ? 1 : slot); s++)
    if (f->pending[s]) { /* stream-side wait: later work on the context's stream sees the gathered table */
      CUDA_TRY(f->ctx, cudaStreamWaitEvent(f->ctx->stream, f->done[s], 0));
      f->pending[s] = false;
    }
  return B200NAV_OK;
}

/* Peer push: this rank has finished reading slot's table of the current epoch (everything enqueued so far on the
 * context's stream).  Writers wait for it before they store the slot's next cycle. */
static int fleet_release_slot(b200nav_fleet* f, int slot) {
  if (!f->push || f->released[slot] >= f->epoch[slot]) return B200NAV_OK;
  FleetAck a;
  memset(&a, 0, sizeof(a));
  for (int p = 0; p < f->world; p++) a.acks[p] = f->acks(p, slot);
  a.world = f->world;
  a.rank = f->rank;
  a.epoch = f->epoch[slot];
  fleet_release_kernel<<<1, 32, 0, f->ctx->stream>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err(f->ctx, B200NAV_ECUDA, "fleet_release_kernel: %s", cudaGetErrorString(e));
  f->released[slot] = f->epoch[slot];
  return B200NAV_OK;
}

int b200nav_fleet_release(b200nav_fleet* f, int slot) {
  if (!f || slot > 1) return B200NAV_EINVAL;
  for (int s = (slot < 0 ? 0 : slot); s <= (slot < 0 ? 1 : slot); s++) {
    const int rc = fleet_release_slot(f, s);
    if (rc) return rc;
  }
  return B200NAV_OK;
}

int b200nav_fleet_push_region(b200nav_fleet* f, int n_local, int n_total, int row0, uint8_t* handle64) {
  if (!f || !handle64 || n_local < 1 || n_total < n_local || row0 < 0 || row0 + n_local > n_total ||
      f->world > B200NAV_MAX_PEERS)
    return B200NAV_EINVAL;
  b200nav_ctx* ctx = f->ctx;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  f->n_local = n_local;
  f->n_total = n_total;
  f->row0 = row0;
  f->region_bytes = sizeof(b200nav_command) * 2 * (size_t)n_total + sizeof(unsigned long long) * 4 * (size_t)f->world + 16;
  CUDA_TRY(ctx, cudaMalloc((void**)&f->region, f->region_bytes));
  CUDA_TRY(ctx, cudaMemset(f->region, 0, f->region_bytes));
  cudaIpcMemHandle_t h;
  CUDA_TRY(ctx, cudaIpcGetMemHandle(&h, f->region));
  static_assert(sizeof(h) == 64, "IPC handle size");
  memcpy(handle64, &h, 64);
  return B200NAV_OK;
}

int b200nav_fleet_push_connect(b200nav_fleet* f, const uint8_t* handles) {
  if (!f || !handles || !f->region) return B200NAV_EINVAL;
  b200nav_ctx* ctx = f->ctx;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  for (int p = 0; p < f->world; p++) {
    if (p == f->rank) {
      f->peer_region[p] = f->region;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + 64 * (size_t)p, 64);
    void* ptr = nullptr;
    CUDA_TRY(ctx, cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    f->peer_region[p] = static_cast<uint8_t*>(ptr);
  }
  f->push = true;
  return B200NAV_OK;
}

void* b200nav_fleet_table(b200nav_fleet* f, int slot) {
  if (!f || !f->region || slot < 0 || slot > 1) return nullptr;
  return reinterpret_cast<b200nav_command*>(f->region) + (size_t)slot * f->n_total;
}

int b200nav_vfh_update_batched_dev_push(b200nav_vfh* v, b200nav_grid* g, const char* layer,
                                        const b200nav_vfh_input* dev_in, b200nav_fleet* f, int slot) {
  if (!v || !g || !dev_in || !f || slot < 0 || slot > 1) return B200NAV_EINVAL;
  if (!f->push) return set_err(v->ctx, B200NAV_EINVAL, "fleet has no peer mappings (b200nav_fleet_push_connect)");
  if (g->n_robots != v->n_robots || v->n_robots != f->n_local)
    return set_err(v->ctx, B200NAV_EINVAL, "grid, vfh and fleet robot counts differ");
  Layer* l = find_layer(g, layer);
  if (!l) return set_err(v->ctx, B200NAV_ENOLAYER, "no layer '%s'", layer ? layer : "(null)");
  b200nav_ctx* ctx = v->ctx;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  { int jrc = join_side(ctx); if (jrc) return jrc; }
  CUDA_TRY(ctx, v->out_buf.reserve(sizeof(b200nav_command) * (size_t)v->n_robots));
  VfhPush push;
  memset(&push, 0, sizeof(push));
  for (int p = 0; p < f->world; p++) {
    push.tables[p] = f->table(p, slot);
    push.flags[p] = f->flags(p, slot);
  }
  push.world = f->world;
  push.rank = f->rank;
  push.row0 = f->row0;
  if (f->epoch[slot] > 0) {
    /* flow control: the slot's previous cycle must have been consumed by EVERY rank before its rows are overwritten.
     * A caller that did not release the slot itself releases it here (its own reads are stream-ordered before). */
    int rrc = fleet_release_slot(f, slot);
    if (rrc) return rrc;
    fleet_wait_acks_kernel<<<1, 32, 0, ctx->stream>>>(f->acks(f->rank, slot), f->world, f->epoch[slot], f->errword());
    cudaError_t we = cudaGetLastError();
    if (we != cudaSuccess) return set_err(ctx, B200NAV_ECUDA, "fleet_wait_acks_kernel: %s", cudaGetErrorString(we));
  }
  push.epoch = ++f->epoch[slot];
  f->push_pending[slot] = true;
  int rc = vfh_launch(v, g, l, dev_in, nullptr, static_cast<b200nav_command*>(v->out_buf.p), 0, v->n_robots, nullptr, &push);
  if (rc) return rc;
  /* the epoch goes out on the fleet's own stream, ordered after the update: off the main stream's critical path */
  CUDA_TRY(ctx, cudaEventRecord(f->ready[slot], ctx->stream));
  CUDA_TRY(ctx, cudaStreamWaitEvent(f->stream, f->ready[slot], 0));
  fleet_flag_kernel<<<1, 32, 0, f->stream>>>(push);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err(ctx, B200NAV_ECUDA, "fleet_flag_kernel: %s", cudaGetErrorString(e));
  return B200NAV_OK;
}

int b200nav_fleet_status(b200nav_fleet* f) {
  if (!f) return B200NAV_EINVAL;
  if (!f->region) return B200NAV_OK;
  b200nav_ctx* ctx = f->ctx;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  CUDA_TRY(ctx, sync_raw(ctx));
  CUDA_TRY(ctx, cudaStreamSynchronize(f->stream));
  int err = 0;
  CUDA_TRY(ctx, cudaMemcpy(&err, f->errword(), sizeof(int), cudaMemcpyDeviceToHost));
  if (err) {
    cudaMemset(f->errword(), 0, sizeof(int));
    return set_err(ctx, B200NAV_ERANGE, "peer push: a rank did not publish its commands within the wait bound "
                                        "(the table of that cycle is incomplete)");
  }
  return B200NAV_OK;
}

int b200nav_fleet_destroy(b200nav_fleet* f) {
  if (!f) return B200NAV_OK;
  cudaSetDevice(f->ctx->device);
  sync_raw(f->ctx);
  for (int p = 0; p < f->world; p++)
    if (p != f->rank && f->peer_region[p]) cudaIpcCloseMemHandle(f->peer_region[p]);
  if (f->region) cudaFree(f->region);
  if (f->stream) cudaStreamSynchronize(f->stream);
  for (int i = 0; i < 2; i++) {
    f->cyc_local[i].release();
    f->cyc_table[i].release();
  }
  if (f->comm && f->CommDestroy) f->CommDestroy(f->comm);
  for (int i = 0; i < 2; i++) {
    if (f->ready[i]) cudaEventDestroy(f->ready[i]);
    if (f->done[i]) cudaEventDestroy(f->done[i]);
  }
  if (f->stream) cudaStreamDestroy(f->stream);
  delete f;
  return B200NAV_OK;
}

/* ================================================================================================================
 * HIMM
 * ============================================================================================================== */

int b200nav_himm_update(b200nav_grid* g, int robot, const char* layer, const b200nav_sample* host_samples, int n,
                        double* bbox) {
  if (!g || robot < 0 || robot >= g->n_robots || n < 0 || (n > 0 && !host_samples)) return B200NAV_EINVAL;
  Layer* l = find_layer(g, layer);
  if (!l) return set_err(g->ctx, B200NAV_ENOLAYER, "no layer '%s'", layer ? layer : "(null)");
  if (n == 0) return B200NAV_OK;
  b200nav_ctx* ctx = g->ctx;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  CUDA_TRY(ctx, g->samples.reserve(sizeof(b200nav_sample) * (size_t)n));
  CUDA_TRY(ctx, cudaMemcpyAsync(g->samples.p, host_samples, sizeof(b200nav_sample) * (size_t)n,
                                cudaMemcpyHostToDevice, ctx->stream));
  /* The binning scratch grows with (samples per robot / 2048) x tiles: very long single-robot batches are applied
   * as several launches over consecutive sample ranges (the order is preserved: launches run back to back). */
  const size_t tiles = grid_tiles(g);
  const size_t bytes_per_chunk = tiles * HIMM_MASK_WORDS * sizeof(uint32_t);
  static const char* budget_env = getenv("B200NAV_MASK_BUDGET_MB"); /* testing aid; default 256 MiB */
  const size_t budget = (size_t)(budget_env ? std::max(1, atoi(budget_env)) : 256) << 20;
  const int max_chunks = (int)std::min<size_t>(4096, std::max<size_t>(1, budget / std::max<size_t>(bytes_per_chunk, 1)));
  const int max_n = max_chunks * HIMM_CHUNK;
  for (int first = 0; first < n; first += max_n) {
    const int cnt = std::min(max_n, n - first);
    int rc = himm_launch(g, l, static_cast<const b200nav_sample*>(g->samples.p) + first, nullptr, robot, 1, cnt, cnt, cnt);
    if (rc) return rc;
  }
  if (bbox) host_touch(host_samples, n, bbox); /* overlaps the kernels */
  return sync_stream(ctx);
}

int b200nav_himm_update_batched(b200nav_grid* g, const char* layer, const b200nav_sample* host_samples,
                                const int32_t* host_offsets, double* bbox) {
  if (!g || !host_offsets) return B200NAV_EINVAL;
  Layer* l = find_layer(g, layer);
  if (!l) return set_err(g->ctx, B200NAV_ENOLAYER, "no layer '%s'", layer ? layer : "(null)");
  b200nav_ctx* ctx = g->ctx;
  const int nr = g->n_robots;
  if (host_offsets[0] != 0) return set_err(ctx, B200NAV_EINVAL, "offsets[0] must be 0");
  for (int r = 0; r < nr; r++)
    if (host_offsets[r + 1] < host_offsets[r]) return set_err(ctx, B200NAV_EINVAL, "offsets must be non-decreasing");
  const int total = host_offsets[nr];
  if (total == 0) return B200NAV_OK;
  if (!host_samples) return B200NAV_EINVAL;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  CUDA_TRY(ctx, g->samples.reserve(sizeof(b200nav_sample) * (size_t)total));
  CUDA_TRY(ctx, g->offsets.reserve(sizeof(int32_t) * (size_t)(nr + 1)));
  CUDA_TRY(ctx, cudaMemcpyAsync(g->samples.p, host_samples, sizeof(b200nav_sample) * (size_t)total,
                                cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(g->offsets.p, host_offsets, sizeof(int32_t) * (size_t)(nr + 1),
                                cudaMemcpyHostToDevice, ctx->stream));
  int max_per = 0;
  for (int r = 0; r < nr; r++) max_per = std::max(max_per, host_offsets[r + 1] - host_offsets[r]);
  int rc = himm_launch(g, l, static_cast<const b200nav_sample*>(g->samples.p),
                       static_cast<const int32_t*>(g->offsets.p), 0, nr, -1, total, max_per);
  if (rc) return rc;
  if (bbox)
    for (int r = 0; r < nr; r++)
      host_touch(host_samples + host_offsets[r], host_offsets[r + 1] - host_offsets[r], bbox + 4 * r);
  return sync_stream(ctx);
}

int b200nav_himm_update_batched_dev(b200nav_grid* g, const char* layer, const b200nav_sample* dev_samples,
                                    const int32_t* dev_offsets, int total, int max_samples_per_robot) {
  if (!g || !dev_offsets || total < 0 || (total > 0 && !dev_samples) || max_samples_per_robot < 0)
    return B200NAV_EINVAL;
  Layer* l = find_layer(g, layer);
  if (!l) return set_err(g->ctx, B200NAV_ENOLAYER, "no layer '%s'", layer ? layer : "(null)");
  CUDA_TRY(g->ctx, cudaSetDevice(g->ctx->device));
  return himm_launch(g, l, dev_samples, dev_offsets, 0, g->n_robots, -1, total, max_samples_per_robot);
}

static int himm_update_cloud_host(b200nav_grid* g, const char* layer, const double* host_origins,
                                  const float* host_xy, const uint8_t* host_clear_end,
                                  const int32_t* host_offsets, double* bbox, bool wait) {
  if (!g || !host_offsets || !host_origins) return B200NAV_EINVAL;
  Layer* l = find_layer(g, layer);
  if (!l) return set_err(g->ctx, B200NAV_ENOLAYER, "no layer '%s'", layer ? layer : "(null)");
  b200nav_ctx* ctx = g->ctx;
  const int nr = g->n_robots;
  if (host_offsets[0] != 0) return set_err(ctx, B200NAV_EINVAL, "offsets[0] must be 0");
  int max_per = 0;
  for (int r = 0; r < nr; r++) {
    if (host_offsets[r + 1] < host_offsets[r]) return set_err(ctx, B200NAV_EINVAL, "offsets must be non-decreasing");
    max_per = std::max(max_per, host_offsets[r + 1] - host_offsets[r]);
  }
  const int total = host_offsets[nr];
  if (total == 0) return B200NAV_OK;
  if (!host_xy) return B200NAV_EINVAL;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  CUDA_TRY(ctx, g->samples.reserve(sizeof(float) * 2 * (size_t)total));
  const int slot = g->cloud_slot;
  g->cloud_slot ^= 1;
  CUDA_TRY(ctx, g->offs2[slot].reserve(sizeof(int32_t) * (size_t)(nr + 1)));
  CUDA_TRY(ctx, g->orig2[slot].reserve(sizeof(double) * 2 * (size_t)nr));
  CloudIn c;
  c.origins = static_cast<const double*>(g->orig2[slot].p);
  c.xy = static_cast<const float*>(g->samples.p);
  if (host_clear_end) {
    CUDA_TRY(ctx, g->clearbuf.reserve((size_t)total));
    c.clear_end = static_cast<const uint8_t*>(g->clearbuf.p);
  }
  HimmArgs a;
  int rc = himm_setup(g, l, nullptr, static_cast<const int32_t*>(g->offs2[slot].p), 0, nr, -1, total, max_per, c, a);
  if (rc) return rc;
  /* Pipeline: the cloud is copied in groups of robots on a second stream while the prep kernel of the previous
   * group runs; the tile kernel starts when everything is binned. */
  const bool piped = total >= (1 << 16) && nr >= 8;
  /* synchronous call: 4 groups hide most of the copy behind the binning of the previous group; asynchronous call:
   * the whole copy already overlaps the previous cycle's kernels, one group = one binning launch is cheaper */
  const int groups = piped ? (wait ? 4 : 1) : 1;
  if (piped && !ctx->copy_stream) {
    CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    ctx->copy_events.resize(8);
    for (auto& e : ctx->copy_events) CUDA_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  cudaStream_t cs = piped ? ctx->copy_stream : ctx->stream;
  /* offsets / origins go first and only wait for the tile kernel that last read this slot (two updates ago): they
   * never queue behind the big cloud copy of a later cycle on the copy engine */
  if (piped && g->tile_done_set[slot]) CUDA_TRY(ctx, cudaStreamWaitEvent(cs, g->tile_done[slot], 0));
  CUDA_TRY(ctx, cudaMemcpyAsync(g->offs2[slot].p, host_offsets, sizeof(int32_t) * (size_t)(nr + 1), cudaMemcpyHostToDevice, cs));
  CUDA_TRY(ctx, cudaMemcpyAsync(g->orig2[slot].p, host_origins, sizeof(double) * 2 * (size_t)nr, cudaMemcpyHostToDevice, cs));
  if (piped) {
    /* The copy stream must not overwrite staging buffers still read by earlier work.  After an asynchronous cloud
     * update the only reader is that update's last binning kernel (event 6): this update's copies then overlap the
     * previous update's tile kernel and whatever follows it on the main stream.  Otherwise: everything enqueued so
     * far. */
    if (!ctx->prep_done_valid) CUDA_TRY(ctx, cudaEventRecord(ctx->copy_events[6], ctx->stream));
    CUDA_TRY(ctx, cudaStreamWaitEvent(cs, ctx->copy_events[6], 0));
  }
  ctx->prep_done_valid = false;
  for (int gi = 0; gi < groups; gi++) {
    const int r_lo = (int)((long long)nr * gi / groups), r_hi = (int)((long long)nr * (gi + 1) / groups);
    const int b_lo = host_offsets[r_lo], b_hi = host_offsets[r_hi];
    if (b_hi > b_lo) {
      CUDA_TRY(ctx, cudaMemcpyAsync(static_cast<float*>(g->samples.p) + 2 * (size_t)b_lo, host_xy + 2 * (size_t)b_lo,
                                    sizeof(float) * 2 * (size_t)(b_hi - b_lo), cudaMemcpyHostToDevice, cs));
      if (host_clear_end)
        CUDA_TRY(ctx, cudaMemcpyAsync(static_cast<uint8_t*>(g->clearbuf.p) + b_lo, host_clear_end + b_lo,
                                      (size_t)(b_hi - b_lo), cudaMemcpyHostToDevice, cs));
    }
    if (piped) {
      CUDA_TRY(ctx, cudaEventRecord(ctx->copy_events[gi], cs));
      CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->copy_events[gi], 0));
    }
    rc = himm_launch_prep(g, a, b_lo, b_hi, r_lo, r_hi);
    if (rc) return rc;
  }
  if (piped && !wait) {
    CUDA_TRY(ctx, cudaEventRecord(ctx->copy_events[6], ctx->stream));
    ctx->prep_done_valid = true;
  }
  rc = himm_launch_tile(g, a);
  if (rc) return rc;
  if (piped) {
    if (!g->tile_done[slot]) CUDA_TRY(ctx, cudaEventCreateWithFlags(&g->tile_done[slot], cudaEventDisableTiming));
    CUDA_TRY(ctx, cudaEventRecord(g->tile_done[slot], ctx->stream));
    g->tile_done_set[slot] = true;
  }
  if (bbox)
    for (int r = 0; r < nr; r++) {
      double* bb = bbox + 4 * r;
      for (int i = host_offsets[r]; i < host_offsets[r + 1]; i++) { /* MapUpdater::touch on start and end */
        const double sx = host_origins[2 * r], sy = host_origins[2 * r + 1];
        const double ex = (double)host_xy[2 * i], ey = (double)host_xy[2 * i + 1];
        bb[0] = std::min(std::min(bb[0], sx), ex);
        bb[1] = std::min(std::min(bb[1], sy), ey);
        bb[2] = std::max(std::max(bb[2], sx), ex);
        bb[3] = std::max(std::max(bb[3], sy), ey);
      }
    }
  return wait ? sync_stream(ctx) : B200NAV_OK;
}

int b200nav_himm_update_cloud_batched(b200nav_grid* g, const char* layer, const double* host_origins,
                                      const float* host_xy, const uint8_t* host_clear_end,
                                      const int32_t* host_offsets, double* bbox) {
  return himm_update_cloud_host(g, layer, host_origins, host_xy, host_clear_end, host_offsets, bbox, true);
}

int b200nav_himm_update_cloud_batched_async(b200nav_grid* g, const char* layer, const double* host_origins,
                                            const float* host_xy, const uint8_t* host_clear_end,
                                            const int32_t* host_offsets) {
  return himm_update_cloud_host(g, layer, host_origins, host_xy, host_clear_end, host_offsets, nullptr, false);
}

int b200nav_himm_update_cloud_batched_dev(b200nav_grid* g, const char* layer, const double* dev_origins,
                                          const float* dev_xy, const uint8_t* dev_clear_end,
                                          const int32_t* dev_offsets, int total, int max_samples_per_robot) {
  if (!g || !dev_offsets || !dev_origins || total < 0 || (total > 0 && !dev_xy) || max_samples_per_robot < 0)
    return B200NAV_EINVAL;
  Layer* l = find_layer(g, layer);
  if (!l) return set_err(g->ctx, B200NAV_ENOLAYER, "no layer '%s'", layer ? layer : "(null)");
  CUDA_TRY(g->ctx, cudaSetDevice(g->ctx->device));
  CloudIn c;
  c.origins = dev_origins;
  c.xy = dev_xy;
  c.clear_end = dev_clear_end;
  return himm_launch(g, l, nullptr, dev_offsets, 0, g->n_robots, -1, total, max_samples_per_robot, c);
}

/* simplifyLaserScan (laser_map_updater.cpp:114-144) restated: which ranges of a scan are projected and with which
 * angle increment.  Returns the number of selected ranges (sel may be NULL to only count). */
int b200nav_scan_select(const b200nav_scan_info* info, int32_t* sel, int cap, float* increment_used) {
  if (!info || info->n_ranges < 0) return B200NAV_EINVAL;
  int n = 0;
  float used = info->angle_increment;
  if (info->decimate && info->angle_increment < 0.017f && info->n_ranges > 0) {
    /* ranges[0] first, then every index at which the float accumulator reaches 0.017 (the comparison is made in
     * double, as `increment >= 0.017` promotes the float) */
    if (sel && n < cap) sel[n] = 0;
    n++;
    float increment = 0.0f;
    for (int i = 0; i < info->n_ranges; i++) {
      increment += info->angle_increment;
      if ((double)increment >= 0.017) {
        used = increment;
        increment = 0.0f;
        if (sel && n < cap) sel[n] = i;
        n++;
      }
    }
  } else {
    for (int i = 0; i < info->n_ranges; i++) {
      if (sel && n < cap) sel[n] = i;
      n++;
    }
  }
  if (increment_used) *increment_used = used;
  return n;
}

enum { SCANS_DEV = 0, SCANS_HOST = 1, SCANS_HOST_ASYNC = 2 };

static int himm_update_scans(b200nav_grid* g, const char* layer, const b200nav_scan_info* info, const double* poses,
                             const float* ranges, int mode) {
  const bool host = mode != SCANS_DEV;
  if (!g || !info || !poses || !ranges || info->n_ranges <= 0) return B200NAV_EINVAL;
  Layer* l = find_layer(g, layer);
  if (!l) return set_err(g->ctx, B200NAV_ENOLAYER, "no layer '%s'", layer ? layer : "(null)");
  b200nav_ctx* ctx = g->ctx;
  const int nr = g->n_robots;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (!g->scan_cache_valid || memcmp(&g->scan_info_cached, info, sizeof(*info)) != 0) {
    std::vector<int32_t> sel((size_t)info->n_ranges + 1);
    float used = 0.f;
    const int n_used = b200nav_scan_select(info, sel.data(), (int)sel.size(), &used);
    if (n_used < 0) return n_used;
    if ((long long)n_used * nr > 0x7fffffffLL) return set_err(ctx, B200NAV_ERANGE, "too many samples for one update");
    std::vector<int32_t> offs((size_t)nr + 1);
    for (int r = 0; r <= nr; r++) offs[r] = r * n_used;
    CUDA_TRY(ctx, sync_raw(ctx)); /* an earlier update may still read the cached arrays */
    CUDA_TRY(ctx, g->scan_sel.reserve(sizeof(int32_t) * (size_t)std::max(n_used, 1)));
    CUDA_TRY(ctx, g->scan_offsets.reserve(sizeof(int32_t) * ((size_t)nr + 1)));
    CUDA_TRY(ctx, cudaMemcpy(g->scan_sel.p, sel.data(), sizeof(int32_t) * (size_t)n_used, cudaMemcpyHostToDevice));
    CUDA_TRY(ctx, cudaMemcpy(g->scan_offsets.p, offs.data(), sizeof(int32_t) * offs.size(), cudaMemcpyHostToDevice));
    g->scan_info_cached = *info;
    g->scan_n_used = n_used;
    g->scan_increment_used = used;
    g->scan_cache_valid = true;
  }
  const int n_used = g->scan_n_used;
  if (n_used == 0) return B200NAV_OK;
  CloudIn c;
  const bool piped = mode == SCANS_HOST_ASYNC;
  if (host) {
    CUDA_TRY(ctx, g->scan_poses.reserve(sizeof(double) * 3 * (size_t)nr));
    CUDA_TRY(ctx, g->scan_ranges.reserve(sizeof(float) * (size_t)nr * info->n_ranges));
    cudaStream_t cs = ctx->stream;
    if (piped) {
      /* as in the asynchronous cloud update: the copies go on the copy stream and only wait for the last binning
       * kernel that read the staging buffers, so they overlap the previous cycle's tile / VFH+ kernels */
      if (!ctx->copy_stream) {
        CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        ctx->copy_events.resize(8);
        for (auto& e : ctx->copy_events) CUDA_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      }
      cs = ctx->copy_stream;
      if (!ctx->prep_done_valid) CUDA_TRY(ctx, cudaEventRecord(ctx->copy_events[6], ctx->stream));
      CUDA_TRY(ctx, cudaStreamWaitEvent(cs, ctx->copy_events[6], 0));
    }
    ctx->prep_done_valid = false;
    CUDA_TRY(ctx, cudaMemcpyAsync(g->scan_poses.p, poses, sizeof(double) * 3 * (size_t)nr, cudaMemcpyHostToDevice, cs));
    CUDA_TRY(ctx, cudaMemcpyAsync(g->scan_ranges.p, ranges, sizeof(float) * (size_t)nr * info->n_ranges,
                                  cudaMemcpyHostToDevice, cs));
    if (piped) {
      CUDA_TRY(ctx, cudaEventRecord(ctx->copy_events[0], cs));
      CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->copy_events[0], 0));
    }
    c.scan_poses = static_cast<const double*>(g->scan_poses.p);
    c.scan_ranges = static_cast<const float*>(g->scan_ranges.p);
  } else {
    c.scan_poses = poses;
    c.scan_ranges = ranges;
  }
  c.scan_sel = (n_used == info->n_ranges) ? nullptr : static_cast<const int32_t*>(g->scan_sel.p);
  c.scan.angle_min = info->angle_min;
  c.scan.increment_used = g->scan_increment_used;
  c.scan.range_min = info->range_min;
  c.scan.range_max = info->range_max;
  c.scan.n_ranges = info->n_ranges;
  c.scan.n_used = n_used;
  c.scan.decimated = (info->decimate && info->angle_increment < 0.017f) ? 1 : 0;
  HimmArgs a;
  int rc = himm_setup(g, l, nullptr, static_cast<const int32_t*>(g->scan_offsets.p), 0, nr, -1, nr * n_used, n_used, c, a);
  if (rc) return rc;
  rc = himm_launch_prep(g, a, 0, nr * n_used, 0, nr);
  if (rc) return rc;
  if (piped) {
    CUDA_TRY(ctx, cudaEventRecord(ctx->copy_events[6], ctx->stream));
    ctx->prep_done_valid = true;
  }
  rc = himm_launch_tile(g, a);
  if (rc) return rc;
  return mode == SCANS_HOST ? sync_stream(ctx) : B200NAV_OK;
}

int b200nav_himm_update_scans_batched(b200nav_grid* g, const char* layer, const b200nav_scan_info* info,
                                      const double* host_poses, const float* host_ranges) {
  return himm_update_scans(g, layer, info, host_poses, host_ranges, SCANS_HOST);
}

int b200nav_himm_update_scans_batched_async(b200nav_grid* g, const char* layer, const b200nav_scan_info* info,
                                            const double* host_poses, const float* host_ranges) {
  return himm_update_scans(g, layer, info, host_poses, host_ranges, SCANS_HOST_ASYNC);
}

int b200nav_himm_update_scans_batched_dev(b200nav_grid* g, const char* layer, const b200nav_scan_info* info,
                                          const double* dev_poses, const float* dev_ranges) {
  return himm_update_scans(g, layer, info, dev_poses, dev_ranges, SCANS_DEV);
}

int b200nav_himm_last_stats(b200nav_grid* g, int64_t* out3) {
  if (!g || !out3) return B200NAV_EINVAL;
  b200nav_ctx* ctx = g->ctx;
  out3[0] = out3[1] = out3[2] = 0;
  if (g->last_total <= 0 || !g->segs.p) return B200NAV_OK;
  CUDA_TRY(ctx, g->stats.reserve(3 * sizeof(unsigned long long)));
  CUDA_TRY(ctx, cudaMemsetAsync(g->stats.p, 0, 3 * sizeof(unsigned long long), ctx->stream));
  const int blocks = std::min((g->last_total + 255) / 256, ctx->sm_count * 8);
  himm_stats_kernel<<<blocks, 256, 0, ctx->stream>>>(static_cast<const BeamSeg*>(g->segs.p), g->last_total,
                                                     static_cast<unsigned long long*>(g->stats.p));
  int rc = check_launch(ctx, "himm_stats_kernel");
  if (rc) return rc;
  CUDA_TRY(ctx, cudaMemcpyAsync(out3, g->stats.p, 3 * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  rc = sync_stream(ctx);
  if (rc) return rc;
  return himm_check_error_flag(g);
}

/* ================================================================================================================
 * VFH+
 * ============================================================================================================== */

void b200nav_vfh_default_params(b200nav_vfh_params* p) {
  if (!p) return;
  memset(p, 0, sizeof(*p));
  /* Steerer::initVfh (steerer.cpp:69-121) */
  p->cell_size = 100;
  p->window_diameter = 30;
  p->sector_angle = 5;
  p->safety_dist_0ms = 10;
  p->safety_dist_1ms = 50;
  p->max_speed = 200;
  p->max_speed_narrow_opening = 200;
  p->max_speed_wide_opening = 300;
  p->max_acceleration = 200;
  p->min_turnrate = 40;
  p->max_turnrate_0ms = 40;
  p->max_turnrate_1ms = 40;
  p->min_turn_radius_safety_factor = 1.0;
  p->free_space_cutoff_0ms = 2000000.0;
  p->obs_cutoff_0ms = 4000000.0;
  p->free_space_cutoff_1ms = 2000000.0;
  p->obs_cutoff_1ms = 4000000.0;
  p->weight_desired_dir = 10.0;
  p->weight_current_dir = 1.0;
  p->robot_radius = 178.0;
  p->submap_length = 1.5;
  p->occupied_threshold = 3.0;
}

static int vfh_reset_state(b200nav_vfh* v) {
  /* VFH ctor + Init (vfh.cpp:90-95, 258-262): Hist = OriginHist = 0, Last_Binary_Hist = 1, angles 90, speed 0 */
  b200nav_ctx* ctx = v->ctx;
  const size_t H = v->tab.c.hist_size, n = v->n_robots;
  std::vector<float> ones(H * n, 1.0f);
  CUDA_TRY(ctx, cudaMemset(v->dev.origin_hist, 0, sizeof(float) * H * n));
  CUDA_TRY(ctx, cudaMemset(v->dev.hist, 0, sizeof(float) * H * n));
  CUDA_TRY(ctx, cudaMemcpy(v->dev.last_binary, ones.data(), sizeof(float) * H * n, cudaMemcpyHostToDevice));
  VfhRobotState s0;
  s0.picked = s0.last_picked = s0.desired = 90.f;
  s0.blocked_radius = 0.f; /* uninitialised in the reference; defined as 0 */
  s0.last_chosen_speed = 0;
  s0.max_speed_for_picked = 0;
  std::vector<VfhRobotState> st(n, s0);
  CUDA_TRY(ctx, cudaMemcpy(v->dev.st, st.data(), sizeof(VfhRobotState) * n, cudaMemcpyHostToDevice));
  std::vector<double> r(n * B200NAV_NRANGES, 5000.0);
  CUDA_TRY(ctx, cudaMemcpy(v->dev.ranges, r.data(), sizeof(double) * r.size(), cudaMemcpyHostToDevice));
  return B200NAV_OK;
}

int b200nav_vfh_create(b200nav_ctx* ctx, const b200nav_vfh_params* p, int n_robots, b200nav_vfh** out) {
  if (!ctx || !p || !out || n_robots < 1) return B200NAV_EINVAL;
  *out = nullptr;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  std::unique_ptr<b200nav_vfh> v(new b200nav_vfh());
  v->ctx = ctx;
  v->params = *p;
  v->n_robots = n_robots;
  int rc = vfh_build_tables(*p, v->tab, ctx->err, sizeof(ctx->err));
  if (rc) return rc;
  const VfhTables& t = v->tab;
  if ((size_t)t.c.front_rows * t.c.window > 65535)
    return set_err(ctx, B200NAV_ERANGE, "VFH window too large (%d front cells > 65535)", t.c.front_rows * t.c.window);
  if ((rc = upload_vec(ctx, &v->d_dir, t.dir))) return rc;
  if ((rc = upload_vec(ctx, &v->d_dist, t.dist))) return rc;
  if ((rc = upload_vec(ctx, &v->d_base, t.base))) return rc;
  if ((rc = upload_vec(ctx, &v->d_thr, t.thr))) return rc;
  if ((rc = upload_vec(ctx, &v->d_kidx, t.kidx))) return rc;
  if ((rc = upload_vec(ctx, &v->d_masks, t.masks))) return rc;
  if ((rc = upload_vec(ctx, &v->d_mtr, t.min_turning_radius))) return rc;
  VfhDev& d = v->dev;
  memset(&d, 0, sizeof(d));
  d.c = t.c;
  d.nf = t.c.front_rows * t.c.window;
  d.dir = v->d_dir;
  d.dist = v->d_dist;
  d.base = v->d_base;
  d.thr = v->d_thr;
  d.kidx = v->d_kidx;
  d.masks = v->d_masks;
  d.mtr = v->d_mtr;
  const size_t H = t.c.hist_size;
  CUDA_TRY(ctx, cudaMalloc((void**)&d.origin_hist, sizeof(float) * H * n_robots));
  CUDA_TRY(ctx, cudaMalloc((void**)&d.hist, sizeof(float) * H * n_robots));
  CUDA_TRY(ctx, cudaMalloc((void**)&d.last_binary, sizeof(float) * H * n_robots));
  CUDA_TRY(ctx, cudaMalloc((void**)&d.st, sizeof(VfhRobotState) * n_robots));
  CUDA_TRY(ctx, cudaMalloc((void**)&d.ranges, sizeof(double) * B200NAV_NRANGES * n_robots));
  if ((rc = vfh_reset_state(v.get()))) return rc;
  *out = v.release();
  return B200NAV_OK;
}

int b200nav_vfh_destroy(b200nav_vfh* v) {
  if (!v) return B200NAV_OK;
  cudaSetDevice(v->ctx->device);
  sync_raw(v->ctx);
  cudaFree(v->d_dir);
  cudaFree(v->d_dist);
  cudaFree(v->d_base);
  cudaFree(v->d_thr);
  cudaFree(v->d_kidx);
  cudaFree(v->d_masks);
  cudaFree(v->d_mtr);
  cudaFree(v->dev.origin_hist);
  cudaFree(v->dev.hist);
  cudaFree(v->dev.last_binary);
  cudaFree(v->dev.st);
  cudaFree(v->dev.ranges);
  v->in_buf.release();
  for (int i = 0; i < 2; i++) {
    v->in2[i].release();
    if (v->in_done[i]) cudaEventDestroy(v->in_done[i]);
  }
  if (v->in_ready) cudaEventDestroy(v->in_ready);
  v->out_buf.release();
  v->ranges_buf.release();
  delete v;
  return B200NAV_OK;
}

int b200nav_vfh_set_current_max_speed(b200nav_vfh* v, int max_speed) {
  if (!v || max_speed < 1) return B200NAV_EINVAL;
  CUDA_TRY(v->ctx, sync_raw(v->ctx));
  vfh_build_min_turning_radius(v->tab, v->params, max_speed);
  int rc = upload_vec(v->ctx, &v->d_mtr, v->tab.min_turning_radius);
  if (rc) return rc;
  v->dev.mtr = v->d_mtr;
  v->dev.c = v->tab.c;
  return B200NAV_OK;
}

int b200nav_vfh_hist_size(const b200nav_vfh* v) { return v ? v->tab.c.hist_size : B200NAV_EINVAL; }
int b200nav_vfh_num_tables(const b200nav_vfh* v) { return v ? v->tab.c.num_tables : B200NAV_EINVAL; }
int b200nav_vfh_get_max_turnrate(const b200nav_vfh* v, int speed) {
  return v ? vfh_get_max_turnrate(v->tab.c.max_turnrate_0ms, v->tab.c.max_turnrate_1ms, speed) : B200NAV_EINVAL;
}

static int vfh_run_host(b200nav_vfh* v, b200nav_grid* g, const char* layer, int robot0, int n,
                        const b200nav_vfh_input* host_in, const double* host_ranges, b200nav_command* host_out,
                        bool wait = true) {
  b200nav_ctx* ctx = v->ctx;
  const Layer* lay = nullptr;
  if (g) {
    if (g->ctx != ctx) return set_err(ctx, B200NAV_EINVAL, "grid and vfh belong to different contexts");
    if (robot0 + n > g->n_robots) return set_err(ctx, B200NAV_EINVAL, "robot index beyond the grid's robots");
    Layer* l = find_layer(g, layer);
    if (!l) return set_err(ctx, B200NAV_ENOLAYER, "no layer '%s'", layer ? layer : "(null)");
    lay = l;
  }
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  CUDA_TRY(ctx, v->out_buf.reserve(sizeof(b200nav_command) * (size_t)n));
  const b200nav_vfh_input* dev_in = nullptr;
  int slot = -1;
  cudaStream_t run = ctx->stream; /* stream of the kernel and of the command copy */
  if (!wait && ctx->copy_stream) {
    /* Pipelined cycles.  (1) The (small) input copy must not queue behind the next cycle's cloud on the copy engine:
     * it is issued on the copy stream as early as the slot's previous reader allows.  (2) The kernel and the copy of
     * the commands run on the side stream, after everything enqueued on the main stream so far (the tile kernel of
     * this cycle): the next cycle's L2 traffic, copies and binning kernel overlap them, and the next tile kernel
     * (the next writer of the layer) joins the side stream first (join_side). */
    if (!ctx->side_stream) {
      CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
      CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_to_side, cudaEventDisableTiming));
      CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_side_done, cudaEventDisableTiming));
    }
    run = ctx->side_stream;
    slot = v->in_slot;
    v->in_slot ^= 1;
    CUDA_TRY(ctx, v->in2[slot].reserve(sizeof(b200nav_vfh_input) * (size_t)n));
    if (!v->in_ready) CUDA_TRY(ctx, cudaEventCreateWithFlags(&v->in_ready, cudaEventDisableTiming));
    if (!v->in_done[slot]) CUDA_TRY(ctx, cudaEventCreateWithFlags(&v->in_done[slot], cudaEventDisableTiming));
    if (v->in_done_set[slot]) CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copy_stream, v->in_done[slot], 0));
    CUDA_TRY(ctx, cudaMemcpyAsync(v->in2[slot].p, host_in, sizeof(b200nav_vfh_input) * (size_t)n,
                                  cudaMemcpyHostToDevice, ctx->copy_stream));
    CUDA_TRY(ctx, cudaEventRecord(v->in_ready, ctx->copy_stream));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_to_side, ctx->stream));
    CUDA_TRY(ctx, cudaStreamWaitEvent(run, ctx->ev_to_side, 0));
    CUDA_TRY(ctx, cudaStreamWaitEvent(run, v->in_ready, 0));
    dev_in = static_cast<const b200nav_vfh_input*>(v->in2[slot].p);
  } else {
    int jrc = join_side(ctx);
    if (jrc) return jrc;
    CUDA_TRY(ctx, v->in_buf.reserve(sizeof(b200nav_vfh_input) * (size_t)n));
    CUDA_TRY(ctx, cudaMemcpyAsync(v->in_buf.p, host_in, sizeof(b200nav_vfh_input) * (size_t)n, cudaMemcpyHostToDevice,
                                  ctx->stream));
    dev_in = static_cast<const b200nav_vfh_input*>(v->in_buf.p);
  }
  const double* dr = nullptr;
  if (host_ranges) {
    CUDA_TRY(ctx, v->ranges_buf.reserve(sizeof(double) * 2 * B200NAV_NRANGES * (size_t)n));
    CUDA_TRY(ctx, cudaMemcpyAsync(v->ranges_buf.p, host_ranges, sizeof(double) * 2 * B200NAV_NRANGES * (size_t)n,
                                  cudaMemcpyHostToDevice, ctx->stream));
    dr = static_cast<const double*>(v->ranges_buf.p);
  }
  int rc = vfh_launch(v, g, lay, dev_in, dr, static_cast<b200nav_command*>(v->out_buf.p), robot0, n, run);
  if (rc) return rc;
  if (slot >= 0) {
    CUDA_TRY(ctx, cudaEventRecord(v->in_done[slot], run));
    v->in_done_set[slot] = true;
  }
  if (run != ctx->stream) { /* the layer's reader is done here; the copy-out below need not hold the next tile kernel */
    if (!ctx->ev_side_kernel) CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_side_kernel, cudaEventDisableTiming));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_side_kernel, run));
  }
  CUDA_TRY(ctx, cudaMemcpyAsync(host_out, v->out_buf.p, sizeof(b200nav_command) * (size_t)n, cudaMemcpyDeviceToHost, run));
  if (run != ctx->stream) {
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_side_done, run));
    ctx->side_pending = true;
  }
  return wait ? sync_stream(ctx) : B200NAV_OK;
}

int b200nav_vfh_update_ranges(b200nav_vfh* v, int robot, const double* ranges361x2, const b200nav_vfh_input* in,
                              b200nav_command* out) {
  if (!v || !ranges361x2 || !in || !out || robot < 0 || robot >= v->n_robots) return B200NAV_EINVAL;
  return vfh_run_host(v, nullptr, nullptr, robot, 1, in, ranges361x2, out);
}

int b200nav_vfh_update_grid(b200nav_vfh* v, b200nav_grid* g, const char* layer, int robot,
                            const b200nav_vfh_input* in, b200nav_command* out) {
  if (!v || !g || !in || !out || robot < 0 || robot >= v->n_robots) return B200NAV_EINVAL;
  return vfh_run_host(v, g, layer, robot, 1, in, nullptr, out);
}

int b200nav_vfh_update_batched(b200nav_vfh* v, b200nav_grid* g, const char* layer, const b200nav_vfh_input* host_in,
                               b200nav_command* host_out) {
  if (!v || !g || !host_in || !host_out) return B200NAV_EINVAL;
  if (g->n_robots != v->n_robots) return set_err(v->ctx, B200NAV_EINVAL, "grid and vfh robot counts differ");
  return vfh_run_host(v, g, layer, 0, v->n_robots, host_in, nullptr, host_out);
}

int b200nav_vfh_update_batched_async(b200nav_vfh* v, b200nav_grid* g, const char* layer,
                                     const b200nav_vfh_input* host_in, b200nav_command* host_out) {
  if (!v || !g || !host_in || !host_out) return B200NAV_EINVAL;
  if (g->n_robots != v->n_robots) return set_err(v->ctx, B200NAV_EINVAL, "grid and vfh robot counts differ");
  return vfh_run_host(v, g, layer, 0, v->n_robots, host_in, nullptr, host_out, false);
}

int b200nav_vfh_update_batched_dev(b200nav_vfh* v, b200nav_grid* g, const char* layer,
                                   const b200nav_vfh_input* dev_in, b200nav_command* dev_out) {
  if (!v || !g || !dev_in || !dev_out) return B200NAV_EINVAL;
  if (g->n_robots != v->n_robots) return set_err(v->ctx, B200NAV_EINVAL, "grid and vfh robot counts differ");
  Layer* l = find_layer(g, layer);
  if (!l) return set_err(v->ctx, B200NAV_ENOLAYER, "no layer '%s'", layer ? layer : "(null)");
  CUDA_TRY(v->ctx, cudaSetDevice(v->ctx->device));
  { int jrc = join_side(v->ctx); if (jrc) return jrc; }
  return vfh_launch(v, g, l, dev_in, nullptr, dev_out, 0, v->n_robots);
}

/* One fleet cycle's steering step from HOST buffers, everything enqueued and nothing waited for: the inputs go up on
 * the copy stream, the VFH+ kernel runs on the side stream behind this cycle's tile kernel and writes this rank's rows,
 * the all-gather and the copy of the gathered table into pinned host memory run on the fleet's stream.  The next
 * cycle's copies, binning and tile kernels overlap all of it (the next tile kernel only joins the side stream). */
int b200nav_fleet_cycle_async(b200nav_fleet* f, b200nav_vfh* v, b200nav_grid* g, const char* layer,
                              const b200nav_vfh_input* host_in, int slot, b200nav_command* host_table) {
  if (!f || !v || !g || !host_in || !host_table || slot < 0 || slot > 1) return B200NAV_EINVAL;
  b200nav_ctx* ctx = v->ctx;
  if (f->ctx != ctx || g->ctx != ctx) return set_err(ctx, B200NAV_EINVAL, "fleet, grid and vfh belong to different contexts");
  if (g->n_robots != v->n_robots) return set_err(ctx, B200NAV_EINVAL, "grid and vfh robot counts differ");
  Layer* lay = find_layer(g, layer);
  if (!lay) return set_err(ctx, B200NAV_ENOLAYER, "no layer '%s'", layer ? layer : "(null)");
  const int n = v->n_robots;
  const size_t row_bytes = sizeof(b200nav_command) * (size_t)n;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (!ctx->copy_stream) {
    CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    ctx->copy_events.resize(8);
    for (auto& e : ctx->copy_events) CUDA_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  if (!ctx->side_stream) {
    CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
    CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_to_side, cudaEventDisableTiming));
    CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_side_done, cudaEventDisableTiming));
  }
  if (f->cyc_local[slot].cap < row_bytes || f->cyc_table[slot].cap < row_bytes * (size_t)f->world) {
    CUDA_TRY(ctx, sync_raw(ctx));
    CUDA_TRY(ctx, cudaStreamSynchronize(f->stream));
    CUDA_TRY(ctx, f->cyc_local[slot].reserve(row_bytes));
    CUDA_TRY(ctx, f->cyc_table[slot].reserve(row_bytes * (size_t)f->world));
  }
  cudaStream_t run = ctx->side_stream;
  /* inputs: double-buffered on the copy stream, as in the single-rank asynchronous update */
  const int in_slot = v->in_slot;
  v->in_slot ^= 1;
  CUDA_TRY(ctx, v->in2[in_slot].reserve(sizeof(b200nav_vfh_input) * (size_t)n));
  if (!v->in_ready) CUDA_TRY(ctx, cudaEventCreateWithFlags(&v->in_ready, cudaEventDisableTiming));
  if (!v->in_done[in_slot]) CUDA_TRY(ctx, cudaEventCreateWithFlags(&v->in_done[in_slot], cudaEventDisableTiming));
  if (v->in_done_set[in_slot]) CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copy_stream, v->in_done[in_slot], 0));
  CUDA_TRY(ctx, cudaMemcpyAsync(v->in2[in_slot].p, host_in, sizeof(b200nav_vfh_input) * (size_t)n, cudaMemcpyHostToDevice,
                                ctx->copy_stream));
  CUDA_TRY(ctx, cudaEventRecord(v->in_ready, ctx->copy_stream));
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev_to_side, ctx->stream));
  CUDA_TRY(ctx, cudaStreamWaitEvent(run, ctx->ev_to_side, 0));
  CUDA_TRY(ctx, cudaStreamWaitEvent(run, v->in_ready, 0));
  /* the gather that last read this slot's rows (two cycles ago) must be done before they are rewritten */
  if (f->cyc_pending[slot]) CUDA_TRY(ctx, cudaStreamWaitEvent(run, f->done[slot], 0));
  int rc = vfh_launch(v, g, lay, static_cast<const b200nav_vfh_input*>(v->in2[in_slot].p), nullptr,
                      static_cast<b200nav_command*>(f->cyc_local[slot].p), 0, n, run);
  if (rc) return rc;
  CUDA_TRY(ctx, cudaEventRecord(v->in_done[in_slot], run));
  v->in_done_set[in_slot] = true;
  if (!ctx->ev_side_kernel) CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_side_kernel, cudaEventDisableTiming));
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev_side_kernel, run));
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev_side_done, run));
  ctx->side_pending = true;
  /* exchange + read-back on the fleet's stream */
  CUDA_TRY(ctx, cudaEventRecord(f->ready[slot], run));
  CUDA_TRY(ctx, cudaStreamWaitEvent(f->stream, f->ready[slot], 0));
  if (f->world > 1) {
    const int r = f->AllGather(f->cyc_local[slot].p, f->cyc_table[slot].p, row_bytes, /*ncclUint8*/ 1, f->comm, f->stream);
    if (r != 0) return set_err(ctx, B200NAV_ECUDA, "ncclAllGather: %s", f->ErrorString ? f->ErrorString(r) : "?");
  }
  CUDA_TRY(ctx, cudaMemcpyAsync(host_table, f->world > 1 ? f->cyc_table[slot].p : f->cyc_local[slot].p,
                                row_bytes * (size_t)f->world, cudaMemcpyDeviceToHost, f->stream));
  CUDA_TRY(ctx, cudaEventRecord(f->done[slot], f->stream));
  f->cyc_pending[slot] = true;
  return B200NAV_OK;
}

/* Blocks the calling thread until the host table of the slot's last b200nav_fleet_cycle_async is complete. */
int b200nav_fleet_cycle_wait(b200nav_fleet* f, int slot) {
  if (!f || slot < 0 || slot > 1) return B200NAV_EINVAL;
  if (f->cyc_pending[slot]) CUDA_TRY(f->ctx, cudaEventSynchronize(f->done[slot]));
  return B200NAV_OK;
}

int b200nav_vfh_read_state(b200nav_vfh* v, int robot, float* origin_hist, float* hist, float* last_binary,
                           float* scalars, int32_t* ints) {
  if (!v || robot < 0 || robot >= v->n_robots) return B200NAV_EINVAL;
  b200nav_ctx* ctx = v->ctx;
  const size_t H = v->tab.c.hist_size;
  CUDA_TRY(ctx, sync_raw(ctx));
  if (origin_hist)
    CUDA_TRY(ctx, cudaMemcpy(origin_hist, v->dev.origin_hist + H * robot, sizeof(float) * H, cudaMemcpyDeviceToHost));
  if (hist) CUDA_TRY(ctx, cudaMemcpy(hist, v->dev.hist + H * robot, sizeof(float) * H, cudaMemcpyDeviceToHost));
  if (last_binary)
    CUDA_TRY(ctx, cudaMemcpy(last_binary, v->dev.last_binary + H * robot, sizeof(float) * H, cudaMemcpyDeviceToHost));
  if (scalars || ints) {
    VfhRobotState s;
    CUDA_TRY(ctx, cudaMemcpy(&s, v->dev.st + robot, sizeof(s), cudaMemcpyDeviceToHost));
    if (scalars) {
      scalars[0] = s.picked;
      scalars[1] = s.last_picked;
      scalars[2] = s.desired;
      scalars[3] = s.blocked_radius;
    }
    if (ints) {
      ints[0] = s.last_chosen_speed;
      ints[1] = s.max_speed_for_picked;
    }
  }
  return B200NAV_OK;
}

int b200nav_vfh_read_ranges(b200nav_vfh* v, int robot, double* ranges361x2) {
  if (!v || !ranges361x2 || robot < 0 || robot >= v->n_robots) return B200NAV_EINVAL;
  double tmp[B200NAV_NRANGES];
  CUDA_TRY(v->ctx, sync_raw(v->ctx));
  CUDA_TRY(v->ctx, cudaMemcpy(tmp, v->dev.ranges + (size_t)B200NAV_NRANGES * robot, sizeof(tmp), cudaMemcpyDeviceToHost));
  for (int i = 0; i < B200NAV_NRANGES; i++) {
    ranges361x2[2 * i] = tmp[i];
    ranges361x2[2 * i + 1] = 0.0;
  }
  return B200NAV_OK;
}

int b200nav_vfh_get_tables(const b200nav_vfh* v, int table, float* dir, float* dist, float* base_mag,
                           uint32_t* sector_masks, int32_t* min_turning_radius) {
  if (!v || table < 0 || table >= v->tab.c.num_tables) return B200NAV_EINVAL;
  const VfhTables& t = v->tab;
  const size_t ww = (size_t)t.c.window * t.c.window;
  if (dir) memcpy(dir, t.dir_xy.data(), ww * sizeof(float));
  if (dist) memcpy(dist, t.dist_xy.data(), ww * sizeof(float));
  if (base_mag) memcpy(base_mag, t.base_xy.data(), ww * sizeof(float));
  if (sector_masks)
    memcpy(sector_masks, t.masks_xy.data() + (size_t)table * ww * t.c.nwords, ww * t.c.nwords * sizeof(uint32_t));
  if (min_turning_radius)
    memcpy(min_turning_radius, t.min_turning_radius.data(), t.min_turning_radius.size() * sizeof(int32_t));
  return B200NAV_OK;
}

/* Statistics hook (not part of the drop-in surface): out[0] = tile work items skipped by the free-space shortcut,
 * out[1] = processed, since the last call. */
int b200nav_himm_debug_tile_stats(b200nav_grid* g, int64_t* out2) {
  if (!g || !out2) return B200NAV_EINVAL;
  out2[0] = out2[1] = 0;
  if (!g->counters.p) return B200NAV_OK;
  int c[8];
  CUDA_TRY(g->ctx, sync_raw(g->ctx));
  CUDA_TRY(g->ctx, cudaMemcpy(c, g->counters.p, sizeof(c), cudaMemcpyDeviceToHost));
  out2[0] = c[4];
  out2[1] = c[5];
  CUDA_TRY(g->ctx, cudaMemset(static_cast<int*>(g->counters.p) + 4, 0, 2 * sizeof(int)));
  return B200NAV_OK;
}

/* Test hook (not part of the drop-in surface): the tile kernel's per-tile summaries of one robot's layer - 64 bits per
 * 64 x 64 tile, bit bc * 8 + br set = the 8 x 8 cells of block column bc / block row br are all known free (value 0).
 * out = tiles_r * tiles_c words (tile index = tile_col * tiles_r + tile_row); all zero for FLOAT layers. */
int b200nav_himm_debug_free_summary(b200nav_grid* g, const char* layer, int robot, uint64_t* out, int cap) {
  if (!g || !out || robot < 0 || robot >= g->n_robots) return B200NAV_EINVAL;
  Layer* l = find_layer(g, layer);
  if (!l) return set_err(g->ctx, B200NAV_ENOLAYER, "no layer '%s'", layer ? layer : "(null)");
  const int nt = (int)grid_tiles(g);
  if (cap < nt) return B200NAV_EINVAL;
  memset(out, 0, sizeof(uint64_t) * (size_t)nt);
  if (!l->free_cols || !l->coded) return nt; /* FLOAT layers keep per-column summaries in the same words */
  CUDA_TRY(g->ctx, sync_raw(g->ctx));
  CUDA_TRY(g->ctx, cudaMemcpy(out, l->free_cols + (size_t)nt * robot, sizeof(uint64_t) * (size_t)nt, cudaMemcpyDeviceToHost));
  return nt;
}

/* Statistics hook: out[0] = 32-beam batches the tile kernel set up, out[1] = batches it dropped before the walk because
 * all their segments only re-clear known-free blocks (FreeBlocks, himm_kernels.cuh), since the last call. */
int b200nav_himm_debug_batch_stats(b200nav_grid* g, int64_t* out2) {
  if (!g || !out2) return B200NAV_EINVAL;
  out2[0] = out2[1] = 0;
  if (!g->counters.p) return B200NAV_OK;
  int c[8];
  CUDA_TRY(g->ctx, sync_raw(g->ctx));
  CUDA_TRY(g->ctx, cudaMemcpy(c, g->counters.p, sizeof(c), cudaMemcpyDeviceToHost));
  out2[0] = c[6];
  out2[1] = c[7];
  CUDA_TRY(g->ctx, cudaMemset(static_cast<int*>(g->counters.p) + 6, 0, 2 * sizeof(int)));
  return B200NAV_OK;
}

/* Test hook (not part of the drop-in surface): force the coalesced-load window path instead of TMA. */
#ifdef VFH_STAGE_CLOCKS
int b200nav_vfh_debug_stage_clocks(unsigned long long* out8, int reset) {
  if (out8) cudaMemcpyFromSymbol(out8, g_vfh_clk, sizeof(unsigned long long) * 8);
  if (reset) {
    unsigned long long z[8] = {0};
    cudaMemcpyToSymbol(g_vfh_clk, z, sizeof(z));
  }
  return 0;
}
#endif

int b200nav_vfh_debug_disable_tma(b200nav_vfh* v, int disable) {
  if (!v) return B200NAV_EINVAL;
  v->tma_disabled = disable != 0;
  return B200NAV_OK;
}

} /* extern "C" */

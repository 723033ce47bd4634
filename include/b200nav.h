/*
 * b200nav.h -- C ABI of the B200-native HIMM mapping + VFH+ steering path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Every entry point names
 * the reference interface it replaces (paths relative to the reference root jmloveyj/ros_navigation).
 * The host side of the reference (ROS nodes, MapProvider, Steerer) keeps calling its own classes; thin
 * C++ shims with the reference's class interfaces (include/b200nav_shim.hpp) forward to these functions.
 *
 * Conventions
 *  - Every function returns B200NAV_OK (0) or a negative B200NAV_E* code; b200nav_last_error() gives text.
 *    No exceptions, no callbacks cross the boundary.  CUDA errors never abort the process.
 *  - The caller owns every buffer it passes.  "host" pointers are ordinary (ideally pinned) host memory,
 *    "dev" pointers are device memory on the context's device.
 *  - Calls on one context are NOT internally locked: serialise them externally, exactly as the reference
 *    serialises map access with MapProvider::mapMutex_ (move_control/src/map_provider.cpp:197).
 *  - All work is enqueued on the context's CUDA stream; entry points that return results to host memory
 *    synchronise that stream before returning, the *_dev variants do not.
 *  - Layers are grid_map::Matrix-compatible: column-major float, rows x cols, NaN = unknown
 *    (grid_map_core/include/grid_map_core/TypeDefs.hpp:16).  With n_robots > 1 a layer is one allocation
 *    [robot][col][row].
 *  - There is no CPU fallback: without a CUDA device b200nav_ctx_create fails with B200NAV_ENODEVICE.
 */
#ifndef B200NAV_H
#define B200NAV_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200NAV_VERSION 100

enum {
  B200NAV_OK = 0,
  B200NAV_EINVAL = -1,    /* bad argument                                  */
  B200NAV_ECUDA = -2,     /* CUDA runtime error (see b200nav_last_error)   */
  B200NAV_ENOMEM = -3,    /* host or device allocation failed              */
  B200NAV_ENOLAYER = -4,  /* unknown layer name                            */
  B200NAV_ERANGE = -5,    /* size outside supported limits                 */
  B200NAV_ENODEVICE = -6  /* no usable CUDA device                         */
};

typedef struct b200nav_ctx b200nav_ctx;
typedef struct b200nav_grid b200nav_grid;
typedef struct b200nav_vfh b200nav_vfh;

/* ------------------------------------------------------------------------------------------------------
 * Context
 * ---------------------------------------------------------------------------------------------------- */

/* device: CUDA ordinal.  cuda_stream: a cudaStream_t to enqueue on (e.g. torch's current stream), or NULL
 * to let the context create its own non-blocking stream. */
int b200nav_ctx_create(int device, void* cuda_stream, b200nav_ctx** out);
int b200nav_ctx_destroy(b200nav_ctx* ctx);
int b200nav_ctx_synchronize(b200nav_ctx* ctx);
/* Pipelining aid for the *_async entry points: fence() marks the current end of the context's stream and returns a
 * ticket; wait() blocks the calling thread until everything enqueued before that fence has completed (tickets may be
 * waited for in any order; only the 8 most recent fences are distinguishable, older ones wait a little longer). */
/* Measurement aid for bench.py: streams write_bytes (stores) and then read_bytes (loads) of scratch memory through
 * L2 on the context's stream so that a following step starts with a cold cache. */
int b200nav_ctx_flush_l2(b200nav_ctx* ctx, size_t write_bytes, size_t read_bytes);
/* Measurement aid for bench.py (SURVEY 8d's second roofline denominator): the rate at which the device retires
 * unordered 4-byte reductions (RED.ADD) at uniformly random words of a scratch buffer of buffer_bytes (rounded down to a
 * power of two; 32 MiB stays in L2, 4 GiB does not).  Synchronous; allocates and frees its scratch. */
int b200nav_ctx_calibrate_red(b200nav_ctx* ctx, size_t buffer_bytes, double* reds_per_second);
int b200nav_ctx_fence(b200nav_ctx* ctx, int* ticket);
int b200nav_ctx_wait(b200nav_ctx* ctx, int ticket);
void* b200nav_ctx_stream(b200nav_ctx* ctx);
/* Last error text of this context (ctx == NULL: of the calling thread's last failed create call). */
const char* b200nav_last_error(b200nav_ctx* ctx);
/* Number of kernels this context launched since creation (bench.py's gpu_launches). */
int64_t b200nav_ctx_launch_count(b200nav_ctx* ctx);
/* Per-kernel device timing with CUDA events on the context's stream (observability; the reference has only a
 * loop-overrun warning, map_provider.cpp:170-173).  enable != 0 starts recording and resets the counters.
 * b200nav_ctx_profile_read synchronises the stream and returns, for kernel `name` ("himm_prep", "himm_tile",
 * "himm_tile_mw", "vfh_update"), the summed device time in milliseconds and the number of timed launches. */
int b200nav_ctx_profile_enable(b200nav_ctx* ctx, int enable);
int b200nav_ctx_profile_read(b200nav_ctx* ctx, const char* name, double* total_ms, int64_t* launches);
/* Restricts the timing to the kernels named in the comma-separated list (NULL or "": all, the default).  Every timed
 * launch costs two event records on the stream, which keep the next kernel's launch from overlapping the previous
 * kernel's tail (measured: 0.016 ms per three-kernel cycle): bench.py times only the dominant kernel inside its timed
 * steps and all of them in a separate pass. */
int b200nav_ctx_profile_select(b200nav_ctx* ctx, const char* names);

/* ------------------------------------------------------------------------------------------------------
 * Grid: device-resident grid_map::GridMap layers for n_robots independent maps of identical size.
 * Replaces the data side of grid_map::GridMap as used by MapProvider (map_provider.cpp:17-41,145-149).
 * ---------------------------------------------------------------------------------------------------- */

/* GridMap::setGeometry (grid_map_core/src/GridMap.cpp:51-70): rows = round(len_x/res), cols = round(len_y/res),
 * length = size*res, startIndex = 0; same position for every robot (change with b200nav_grid_set_geometry). */
int b200nav_grid_create(b200nav_ctx* ctx, double len_x, double len_y, double res, double pos_x, double pos_y,
                        int n_robots, b200nav_grid** out);
int b200nav_grid_destroy(b200nav_grid* grid);
int b200nav_grid_size(const b200nav_grid* grid, int* rows, int* cols, int* n_robots);
/* GridMap::add(layer) with NaN fill (MapUpdater ctor, move_control/include/move_control/map_updater.h:10-14). */
int b200nav_grid_add_layer(b200nav_grid* grid, const char* name);
/* Make `alias` name the same device memory as `target` (zero-copy form of map_["master"] = map_["laser"],
 * map_provider.cpp:221).  */
int b200nav_grid_alias_layer(b200nav_grid* grid, const char* alias, const char* target);
/* map_[dst] = map_[src] for all robots (map_provider.cpp:221, the copying form). */
int b200nav_grid_copy_layer(b200nav_grid* grid, const char* dst, const char* src);
/* The two-layer compose that MapProvider::composeMasterMapFromLayerdMap carries commented out
 * (move_control/src/map_provider.cpp:218-220): dst = (range is NaN and laser is not ? 0 : range) +
 * (laser is NaN and range is not ? 0 : laser), all robots.  The destination becomes a FLOAT-format layer (sums leave
 * the HIMM value set) and must not alias a source.  Asynchronous on the context's stream.  The compose the reference
 * actually runs (master = laser, :221) is b200nav_grid_copy_layer / b200nav_grid_alias_layer. */
int b200nav_grid_compose_master(b200nav_grid* grid, const char* dst, const char* range_layer, const char* laser_layer);
/* GridMap::clear(layer) / clearAll (GridMap.cpp:605-629): NaN fill.  layer == NULL clears every layer. */
int b200nav_grid_clear(b200nav_grid* grid, const char* layer);
/* Whole-layer transfer for one robot; `colmajor` is rows*cols floats laid out like Eigen::MatrixXf::data(). */
int b200nav_grid_upload(b200nav_grid* grid, int robot, const char* layer, const float* colmajor);
int b200nav_grid_download(b200nav_grid* grid, int robot, const char* layer, float* colmajor);
/* position_ / startIndex_ of one robot's map (GridMap::setPosition, GridMap.hpp:499-516). */
int b200nav_grid_set_geometry(b200nav_grid* grid, int robot, double pos_x, double pos_y, int start0, int start1);
int b200nav_grid_get_geometry(const b200nav_grid* grid, int robot, double* pos_x, double* pos_y, int* start0,
                              int* start1);
/* GridMap::move(position) (GridMap.cpp:346-412): circular-buffer shift + NaN-fill of the dropped strips in
 * every layer.  *moved = 1 if the start index changed. */
int b200nav_grid_move(b200nav_grid* grid, int robot, double x, double y, int* moved);
/* GridMapRosConverter::toOccupancyGrid (grid_map_ros/src/GridMapRosConverter.cpp:251-287): int8 [-1,0..100],
 * reversed cell order, unwrapped index.  out_host = rows*cols bytes. */
int b200nav_grid_to_occupancy(b200nav_grid* grid, int robot, const char* layer, float data_min, float data_max,
                              int8_t* out_host);
/* MapGlobalPlanner::ifBlocked (move_control/include/move_control/map_global_planner.h:39-54, grid_map::CircleIterator)
 * for n query points (host_xy: n*2 doubles): out_host[i] = 1 if any cell whose centre lies within `radius` of the
 * point holds a non-NaN value > 0.  The RRT planner (rrt_planner.cpp:53) can test candidates against the
 * device-resident master layer without downloading it. */
int b200nav_grid_query_blocked(b200nav_grid* grid, int robot, const char* layer, const double* host_xy, int n,
                               double radius, uint8_t* out_host);
/* Tell the library that layer `layer` was written through b200nav_grid_layer_devptr (robot < 0: all robots); it drops
 * the cached per-tile knowledge the HIMM kernel keeps about that layer. */
int b200nav_grid_layer_written(b200nav_grid* grid, const char* layer, int robot);
/* Device pointer of a layer ([robot][col][row] floats, the reference's Eigen::MatrixXf layout per robot) for
 * zero-copy consumers.  Layers normally live in HBM as one byte per cell (see b200nav_grid_layer_format); asking for
 * the raw pointer converts the layer to the float layout for good (slower HIMM updates).  NULL if there is no such
 * layer. */
void* b200nav_grid_layer_devptr(b200nav_grid* grid, const char* layer);
/* GridMap::exists (grid_map_core/src/GridMap.cpp:116-119): 1 if the layer exists, else 0. */
int b200nav_grid_has_layer(const b200nav_grid* grid, const char* layer);
/* Device format of a layer.  CODED: one byte per cell in 64 x 64-cell tile records - possible while every value is NaN
 * or one of 0, 10, ..., 180, which is all the HIMM update ever writes (map_updater.h:49-71).  FLOAT: the reference's
 * float matrix; a layer switches to it when b200nav_grid_upload brings any other value or the raw device pointer is
 * requested.  Results are identical in both formats. */
#define B200NAV_LAYER_FLOAT 0
#define B200NAV_LAYER_CODED 1
int b200nav_grid_layer_format(b200nav_grid* grid, const char* layer);

/* ------------------------------------------------------------------------------------------------------
 * Fleet: batched multi-GPU mode (no counterpart in the reference: one robot, one process).  One process per GPU,
 * robots block-partitioned, NO collective on grids or scans; once per cycle the 16-byte b200nav_command records of
 * all robots are all-gathered over NVLink so that every rank sees the whole fleet.  NCCL is bound at run time
 * (libnccl.so.2 via dlopen).  The gather runs on its own stream: start it after the VFH+ update that wrote
 * dev_local, keep launching the next cycle, and call b200nav_fleet_wait(slot) before anything on the context's
 * stream reads dev_table or rewrites dev_local (two slots = double buffering).
 *   id128: 128 bytes from b200nav_fleet_unique_id on rank 0, distributed to all ranks by the launcher.
 * ------------------------------------------------------------------------------------------------------ */
typedef struct b200nav_fleet b200nav_fleet;
int b200nav_fleet_unique_id(uint8_t* id128);
int b200nav_fleet_create(b200nav_ctx* ctx, const uint8_t* id128, int rank, int world, b200nav_fleet** out);
int b200nav_fleet_gather_async(b200nav_fleet* fleet, int slot, const void* dev_local, void* dev_table,
                               size_t bytes_per_rank);
/* slot < 0: both slots. */
int b200nav_fleet_wait(b200nav_fleet* fleet, int slot);
/* Fused exchange ("peer push", optional): instead of a collective after the VFH+ update, the VFH+ kernel itself
 * stores every robot's command into row (row0 + robot) of the table of EVERY rank through NVLink peer mappings and
 * then publishes the cycle's epoch in every rank's flag array; b200nav_fleet_wait(slot) enqueues a small kernel that
 * waits (bounded, about a second) for all ranks' epochs, after which b200nav_fleet_table(slot) holds the whole fleet.
 * Setup: every rank calls _push_region (allocates its table + flags and returns a 64-byte CUDA IPC handle), the
 * launcher all-gathers the handles, every rank calls _push_connect with all of them (rank order).  Contract as
 * above: b200nav_fleet_wait(slot) before the update that writes slot again; at most one cycle per slot in flight.
 * Flow control: after its last read of b200nav_fleet_table(slot) a rank calls b200nav_fleet_release(slot) (enqueued on
 * the context's stream, i.e. ordered after reads enqueued there); a rank's next update of that slot waits until every
 * rank has released it, so a rank that runs ahead can never overwrite a table a slower rank is still reading.  An
 * update whose caller did not release the slot releases it itself (correct, but ranks then run in lock step). */
int b200nav_fleet_push_region(b200nav_fleet* fleet, int n_local, int n_total, int row0, uint8_t* handle64);
int b200nav_fleet_push_connect(b200nav_fleet* fleet, const uint8_t* handles);
void* b200nav_fleet_table(b200nav_fleet* fleet, int slot);
/* slot < 0: both slots.  No-op for the NCCL exchange. */
int b200nav_fleet_release(b200nav_fleet* fleet, int slot);
/* Synchronises and reports (B200NAV_ERANGE) whether a wait of the peer push ran into its bound since the last call. */
int b200nav_fleet_status(b200nav_fleet* fleet);
int b200nav_fleet_destroy(b200nav_fleet* fleet);

/* ------------------------------------------------------------------------------------------------------
 * HIMM update.  Replaces LaserMapUpdater::updateMap / RangeMapUpdater::updateMap
 * (move_control/src/laser_map_updater.cpp:7-21, range_map_updater.cpp:7-21) and MapUpdater::lineOnMap /
 * clearCell / markCell (map_updater.h:38-71) including grid_map::LineIterator
 * (grid_map_core/src/iterators/LineIterator.cpp:16-150).
 * ---------------------------------------------------------------------------------------------------- */

/* MapUpdater::RangeSample (map_updater.h:28-32): map-frame metres. */
typedef struct {
  double sx, sy;      /* start  */
  double ex, ey;      /* end    */
  int32_t clear_end;  /* ifClearEnd: non-zero = do not mark the end cell */
  int32_t reserved;
} b200nav_sample;

/* Apply n samples IN ORDER to layer `layer` of robot `robot`.  bbox = {minX,minY,maxX,maxY}, in/out, updated
 * like MapUpdater::touch (map_updater.h:73-78); may be NULL.  Result is bit-identical to the reference. */
int b200nav_himm_update(b200nav_grid* grid, int robot, const char* layer, const b200nav_sample* host_samples, int n,
                        double* bbox);
/* Batched: samples of robot r are host_samples[offsets[r] .. offsets[r+1]) (offsets has n_robots+1 entries),
 * each robot's samples applied in order to its own map.  bbox: n_robots*4 doubles in/out, or NULL. */
int b200nav_himm_update_batched(b200nav_grid* grid, const char* layer, const b200nav_sample* host_samples,
                                const int32_t* host_offsets, double* bbox);
/* Same with samples/offsets already in device memory; asynchronous (no stream sync).  total = offsets[n_robots];
 * max_samples_per_robot = an upper bound of offsets[r+1]-offsets[r] (sizes the binning scratch; a robot exceeding
 * it is reported as B200NAV_ERANGE by the next b200nav_himm_last_stats call). */
int b200nav_himm_update_batched_dev(b200nav_grid* grid, const char* layer, const b200nav_sample* dev_samples,
                                    const int32_t* dev_offsets, int total, int max_samples_per_robot);

/* Compact "cloud" form of the same update: what LaserMapUpdater::bufferIncomingMsg actually holds per message
 * (laser_map_updater.cpp:46-70) - ONE start point per robot (the laser origin, 2 doubles) and per cloud point the
 * float32 x / y of the PointCloud2 plus the ifClearEnd flag - instead of 40-byte RangeSamples (5x less host->device
 * traffic).  Equivalent to the sample form with sx,sy = origin and ex,ey = (double)x,(double)y, bit for bit.
 *   origins: n_robots*2 doubles; xy: total*2 floats; clear_end: total bytes or NULL (all 0); offsets: n_robots+1. */
int b200nav_himm_update_cloud_batched(b200nav_grid* grid, const char* layer, const double* host_origins,
                                      const float* host_xy, const uint8_t* host_clear_end,
                                      const int32_t* host_offsets, double* bbox);
/* Same, but returns as soon as the copies and kernels are enqueued.  The host arrays must stay valid and unchanged
 * until the work has completed (b200nav_ctx_wait on a later fence, b200nav_ctx_synchronize, or any synchronous call);
 * use pinned memory, else the copies are staged synchronously.  Back-to-back asynchronous cycles overlap: the
 * host->device copy of the next cloud runs while the tile kernel and the VFH+ kernel of the previous cycle execute
 * (it only waits for the previous cycle's binning kernel, the last reader of the staged cloud). */
int b200nav_himm_update_cloud_batched_async(b200nav_grid* grid, const char* layer, const double* host_origins,
                                            const float* host_xy, const uint8_t* host_clear_end,
                                            const int32_t* host_offsets);
/* ------------------------------------------------------------------------------------------------------
 * Scan form: raw LaserScans + the sensor pose, projected on the device.  Replaces the intake side of
 * LaserMapUpdater::bufferIncomingMsg (move_control/src/laser_map_updater.cpp:38-144: getLaserOriginOnGlobal,
 * simplifyLaserScan, laser_geometry's projection, the RangeSample loop).  laser_geometry / tf are not part of the
 * reference tree; their arithmetic is restated as an explicit specification in
 * ros_navigation_b200/csrc/scan_project.h (parity with the reference unpinned there, bit-exact with oracle/).
 *   info:   the LaserScan header fields shared by all robots; decimate != 0 applies simplifyLaserScan (the
 *           reference always does: scans finer than 0.017 rad are thinned to about one reading per degree)
 *   poses:  n_robots * 3 doubles: sensor x, y, yaw in the map frame at the scan's stamp
 *   ranges: n_robots * n_ranges floats
 * 4 bytes per reading cross PCIe instead of the 40-byte RangeSample.
 * ------------------------------------------------------------------------------------------------------ */
typedef struct b200nav_scan_info {
  float angle_min, angle_increment, range_min, range_max;
  int32_t n_ranges;
  int32_t decimate;
} b200nav_scan_info;
/* simplifyLaserScan: indices of the readings that are projected (sel may be NULL) and the angle increment the
 * projection uses; returns their number. */
int b200nav_scan_select(const b200nav_scan_info* info, int32_t* sel, int cap, float* increment_used);
int b200nav_himm_update_scans_batched(b200nav_grid* grid, const char* layer, const b200nav_scan_info* info,
                                      const double* host_poses, const float* host_ranges);
/* Enqueue-only form (pinned host arrays, valid until the work completed: see b200nav_himm_update_cloud_batched_async);
 * the copy of the next cycle's ranges overlaps the current cycle's tile and VFH+ kernels. */
int b200nav_himm_update_scans_batched_async(b200nav_grid* grid, const char* layer, const b200nav_scan_info* info,
                                            const double* host_poses, const float* host_ranges);
/* Same with device arrays; asynchronous. */
int b200nav_himm_update_scans_batched_dev(b200nav_grid* grid, const char* layer, const b200nav_scan_info* info,
                                          const double* dev_poses, const float* dev_ranges);

int b200nav_himm_update_cloud_batched_dev(b200nav_grid* grid, const char* layer, const double* dev_origins,
                                          const float* dev_xy, const uint8_t* dev_clear_end,
                                          const int32_t* dev_offsets, int total, int max_samples_per_robot);

/* Work statistics of the LAST himm update of this grid (all robots of that call): out[0] = cell visits
 * (sum over beams of the Bresenham cell count), out[1] = marks, out[2] = beams.  Used for the algorithmic-byte
 * accounting of the roofline (8 B per visit + 8 B per mark + 36 B per beam). */
int b200nav_himm_last_stats(b200nav_grid* grid, int64_t* out3);

/* ------------------------------------------------------------------------------------------------------
 * VFH+.  Replaces move_control::VFH (move_control/include/move_control/vfh.h:182-361,
 * move_control/src/vfh.cpp) and Steerer::getRangesFromSubmap (move_control/src/steerer.cpp:147-191).
 * ---------------------------------------------------------------------------------------------------- */

/* The 19 VFH constructor arguments (vfh.h:185-203) + SetRobotRadius + the Steerer constants. */
typedef struct {
  double cell_size;                     /* mm                                   */
  int32_t window_diameter;              /* cells                                */
  int32_t sector_angle;                 /* deg                                  */
  double safety_dist_0ms;               /* mm                                   */
  double safety_dist_1ms;
  int32_t max_speed;                    /* mm/s                                 */
  int32_t max_speed_narrow_opening;
  int32_t max_speed_wide_opening;
  int32_t max_acceleration;             /* mm/s^2                               */
  int32_t min_turnrate;                 /* deg/s (unused by the algorithm)      */
  int32_t max_turnrate_0ms;
  int32_t max_turnrate_1ms;
  int32_t reserved0;
  double min_turn_radius_safety_factor;
  double free_space_cutoff_0ms;
  double obs_cutoff_0ms;
  double free_space_cutoff_1ms;
  double obs_cutoff_1ms;
  double weight_desired_dir;
  double weight_current_dir;
  double robot_radius;                  /* mm, VFH::SetRobotRadius              */
  double submap_length;                 /* m, Steerer: Length(1.5,1.5) (steerer.cpp:158) */
  double occupied_threshold;            /* cells with value <= this are free (steerer.cpp:167: 3) */
} b200nav_vfh_params;

/* Steerer::initVfh defaults (steerer.cpp:69-121). */
void b200nav_vfh_default_params(b200nav_vfh_params* p);

/* Per-robot inputs of one decision (Steerer::update, steerer.cpp:221-263). */
typedef struct {
  double x, y, yaw;         /* robot pose in the map frame (MapProvider::getRobotPos)        */
  double dt;                /* seconds since the previous update (reference: gettimeofday)   */
  int32_t current_speed;    /* mm/s                                                          */
  float goal_direction;     /* deg, 0 = robot's right, 90 = ahead                            */
  float goal_distance;      /* mm                                                            */
  float goal_tolerance;     /* mm                                                            */
} b200nav_vfh_input;

/* Per-robot result; the 16-byte record that batched mode all-gathers. */
typedef struct {
  int32_t speed;            /* chosen_speed, mm/s     */
  int32_t turnrate;         /* chosen_turnrate, deg/s */
  float picked_angle;       /* VFH::GetPickedAngle    */
  uint32_t flags;           /* B200NAV_CMD_* bits     */
} b200nav_command;

#define B200NAV_CMD_EMERGENCY 1u   /* obstacle inside the safety distance (vfh.cpp:1020-1034) */
#define B200NAV_CMD_HEMMED_IN 2u   /* no candidate direction (vfh.cpp:720-728)                */
#define B200NAV_CMD_CANT_TURN 4u   /* Cant_Turn_To_Goal (vfh.cpp:612-654)                     */
#define B200NAV_CMD_NO_SUBMAP 8u   /* window could not be formed; treated as no obstacles     */

/* VFH::VFH + SetRobotRadius + Init for n_robots independent controllers sharing one parameter set
 * (vfh.cpp:53-110,237-416). */
int b200nav_vfh_create(b200nav_ctx* ctx, const b200nav_vfh_params* p, int n_robots, b200nav_vfh** out);
int b200nav_vfh_destroy(b200nav_vfh* vfh);
/* VFH::SetCurrentMaxSpeed (vfh.cpp:144-166). */
int b200nav_vfh_set_current_max_speed(b200nav_vfh* vfh, int max_speed);
int b200nav_vfh_hist_size(const b200nav_vfh* vfh);       /* VFH::getHistSize        */
int b200nav_vfh_num_tables(const b200nav_vfh* vfh);      /* NUM_CELL_SECTOR_TABLES  */
int b200nav_vfh_get_max_turnrate(const b200nav_vfh* vfh, int speed); /* VFH::GetMaxTurnrate (vfh.cpp:130-138) */

/* VFH::Update_VFH (vfh.cpp:480-605) for one robot with a host pseudo-scan double[361][2].
 * `in->x,y,yaw` are ignored. */
int b200nav_vfh_update_ranges(b200nav_vfh* vfh, int robot, const double* ranges361x2, const b200nav_vfh_input* in,
                              b200nav_command* out);
/* Steerer::getRangesFromSubmap + VFH::Update_VFH fused, reading layer `layer` of the device grid. */
int b200nav_vfh_update_grid(b200nav_vfh* vfh, b200nav_grid* grid, const char* layer, int robot,
                            const b200nav_vfh_input* in, b200nav_command* out);
/* All robots; host arrays of n_robots entries. */
int b200nav_vfh_update_batched(b200nav_vfh* vfh, b200nav_grid* grid, const char* layer,
                               const b200nav_vfh_input* host_in, b200nav_command* host_out);
/* Same, but returns once the input copy, the kernel and the copy of the commands into host_out are enqueued
 * (host arrays: see b200nav_himm_update_cloud_batched_async). */
int b200nav_vfh_update_batched_async(b200nav_vfh* vfh, b200nav_grid* grid, const char* layer,
                                     const b200nav_vfh_input* host_in, b200nav_command* host_out);
/* Same with device arrays; asynchronous. */
int b200nav_vfh_update_batched_dev(b200nav_vfh* vfh, b200nav_grid* grid, const char* layer,
                                   const b200nav_vfh_input* dev_in, b200nav_command* dev_out);

/* Batched update whose kernel also delivers the commands to every rank (see b200nav_fleet_push_region); device
 * inputs, asynchronous.  The fleet's n_local must equal the robot count of vfh and grid. */
int b200nav_vfh_update_batched_dev_push(b200nav_vfh* vfh, b200nav_grid* grid, const char* layer,
                                        const b200nav_vfh_input* dev_in, b200nav_fleet* fleet, int slot);

/* Steerer::update's caller glue (move_control/src/steerer.cpp:221-258) for n robots, on the HOST (as in the
 * reference; see ros_navigation_b200/csrc/steer_host.cpp for why): for every robot skip the waypoints of its plan
 * that lie within `tolerance` mm (steerer.cpp:234-250, plan_index in/out; Steerer::acceptPlan starts it at 1) and fill
 * inputs[r] with the pose, current_speed = (int)(odom_speed[r] * 1000) (:252,262), goal_distance = hypotf(dx, dy) in mm
 * and goal_direction = RAD2DEG(normalize_angle_positive(atan2f(dy, dx) - yaw + pi/2)) (:254); `dt` is left alone.
 *   poses: n*3 doubles (x, y, yaw: MapProvider::getRobotPos); waypoints: 2 doubles each, robot r owns
 *   [wp_offsets[r], wp_offsets[r+1]); odom_speed: m/s or NULL; plan_done[r] = 1 when the plan is exhausted (the
 *   reference then clears ifPlanReady_ and publishes nothing; inputs[r] gets goal_distance 0).  May be NULL. */
int b200nav_steer_update_goals(int n, const double* poses, const double* waypoints, const int32_t* wp_offsets,
                               int32_t* plan_index, float tolerance, const double* odom_speed,
                               b200nav_vfh_input* inputs, uint8_t* plan_done);

/* One fleet cycle's steering step from pinned HOST buffers, enqueue only (the N > 1 form of
 * b200nav_vfh_update_batched_async): inputs up, VFH+ kernel for this rank's robots, all-gather of the commands, the
 * whole fleet's table (world x n_robots records, rank order) back into host_table - on side streams, overlapping the
 * next cycle's copies and kernels.  Two slots; b200nav_fleet_cycle_wait(slot) blocks until host_table of the slot's
 * last cycle is complete (call it before reusing the slot's host buffers).  NCCL exchange only. */
int b200nav_fleet_cycle_async(b200nav_fleet* fleet, b200nav_vfh* vfh, b200nav_grid* grid, const char* layer,
                              const b200nav_vfh_input* host_in, int slot, b200nav_command* host_table);
int b200nav_fleet_cycle_wait(b200nav_fleet* fleet, int slot);

/* Read back per-robot state after an update (any pointer may be NULL):
 *  origin_hist / hist / last_binary: hist_size floats (VFH::OriginHist, VFH::Hist, Last_Binary_Hist);
 *  scalars[4] = {Picked_Angle, Last_Picked_Angle, Desired_Angle, Blocked_Circle_Radius};
 *  ints[2]    = {last_chosen_speed, Max_Speed_For_Picked_Angle}. */
int b200nav_vfh_read_state(b200nav_vfh* vfh, int robot, float* origin_hist, float* hist, float* last_binary,
                           float* scalars, int32_t* ints);
/* The pseudo-scan (double[361][2], column 1 = 0) the last *_grid update of this robot built. */
int b200nav_vfh_read_ranges(b200nav_vfh* vfh, int robot, double* ranges361x2);
/* Tables built by Init, for parity checks: per cell [x][y] (W*W): direction, distance, base magnitude;
 * sector masks of table `table`: W*W*nwords uint32 (nwords = (hist_size+31)/32); min_turning_radius:
 * current_max_speed+1 ints.  Any pointer may be NULL. */
int b200nav_vfh_get_tables(const b200nav_vfh* vfh, int table, float* dir, float* dist, float* base_mag,
                           uint32_t* sector_masks, int32_t* min_turning_radius);

#ifdef __cplusplus
}
#endif
#endif /* B200NAV_H */

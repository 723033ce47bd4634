/*
 * oracle_api.h -- C ABI of the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT).
 *
 * The oracle restates, in plain dependency-free C++, the reference's CPU algorithm for the
 * HIMM certainty-grid update and the grid-window -> pseudo-scan stage.  Every function cites the
 * reference file:line it follows (paths relative to the reference root).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library, and only as the checker / CPU baseline.  The product path (ros_navigation_b200/csrc)
 * never links or calls it.
 *
 * Parity pinning: the substrate (index math, inside test, line iterator, move) is pinned by the
 * vendored grid_map known-answer tests (tests/test_oracle_kat.py restates them).  The HIMM cell
 * arithmetic and getRangesFromSubmap have NO golden vectors in the reference ("parity unpinned" by
 * reference tests); they are short restatements of map_updater.h:38-71 and steerer.cpp:147-191.
 */
#ifndef B200NAV_ORACLE_API_H
#define B200NAV_ORACLE_API_H

#ifdef __cplusplus
extern "C" {
#endif

/* grid_map::GridMap geometry (GridMap.hpp:493-516): size, resolution, length, position, startIndex. */
typedef struct {
  int rows, cols;        /* size_(0), size_(1)                         */
  double res;            /* resolution_                                */
  double len_x, len_y;   /* length_ = size * resolution                */
  double pos_x, pos_y;   /* position_ (map centre in the map frame)    */
  int start0, start1;    /* startIndex_ (circular buffer origin)       */
} oracle_geom;

/* move_control::MapUpdater::RangeSample (map_updater.h:28-32). */
typedef struct {
  double sx, sy, ex, ey;
  int clear_end;         /* ifClearEnd */
  int pad_;
} oracle_sample;

/* GridMap::setGeometry (GridMap.cpp:51-70). */
void oracle_geom_init(oracle_geom* g, double len_x, double len_y, double res, double pos_x, double pos_y);

/* checkIfPositionWithinMap (GridMapMath.cpp:146-159). */
int oracle_is_inside(const oracle_geom* g, double x, double y);
/* getIndexFromPosition (GridMapMath.cpp:130-144). returns 1 on success. */
int oracle_index_from_position(const oracle_geom* g, double x, double y, int* r, int* c);
/* getPositionFromIndex (GridMapMath.cpp:115-128). */
int oracle_position_from_index(const oracle_geom* g, int r, int c, double* x, double* y);
/* getIndexShiftFromPositionShift (GridMapMath.cpp:170-184). */
void oracle_index_shift_from_position_shift(double dx, double dy, double res, int* s0, int* s1);

/* grid_map::LineIterator(map, start, end) (LineIterator.cpp:16-23,60-75,92-150).
 * Writes up to cap (row,col) pairs; returns the number of cells of the line (0 = no line). */
int oracle_line_cells(const oracle_geom* g, double sx, double sy, double ex, double ey, int* rc, int cap);

/* LaserMapUpdater::updateMap + MapUpdater::lineOnMap/clearCell/markCell
 * (laser_map_updater.cpp:7-21, map_updater.h:38-78).  layer = column-major rows x cols float.
 * bbox = {minX,minY,maxX,maxY} in/out ("touch").  Returns the number of cell visits (clears). */
long long oracle_himm_update(const oracle_geom* g, float* layer, const oracle_sample* s, int n, double* bbox);
/* Same, but with the reference's per-cell cost model: a string-keyed unordered_map lookup for every
 * clearCell/markCell (GridMap.cpp:134-141 via map_updater.h:53,62).  Used by the CPU baseline only. */
long long oracle_himm_update_as_written(const oracle_geom* g, float* layer, const oracle_sample* s, int n,
                                        double* bbox);
/* Count visits + marks without touching a layer (algorithmic-byte accounting for the roofline). */
void oracle_himm_count(const oracle_geom* g, const oracle_sample* s, int n, long long* visits, long long* marks);

/* getSubmapInformation (GridMapMath.cpp:246-296). Returns 1 on success.
 * tl = buffer index of the top-left cell, size = submap size, sub_pos/len = submap geometry. */
int oracle_submap_info(const oracle_geom* g, double cx, double cy, double lx, double ly,
                       int* tl_r, int* tl_c, int* size_r, int* size_c,
                       double* sub_px, double* sub_py, double* sub_lx, double* sub_ly);
/* GridMap::getSubmap data copy (GridMap.cpp:294-339) for one layer: out = size_r x size_c column-major. */
int oracle_get_submap(const oracle_geom* g, const float* layer, double cx, double cy, double lx, double ly,
                      float* out, int out_cap, int* size_r, int* size_c);

/* Steerer::getRangesFromSubmap (steerer.cpp:147-191). ranges = double[361][2] (column 1 untouched). */
void oracle_ranges_from_submap(const oracle_geom* g, const float* master, double rx, double ry, double yaw,
                               double submap_len, double* ranges361x2);

/* GridMap::move (GridMap.cpp:346-412) applied to nlayers layers. Returns 1 if the map moved. */
int oracle_move(oracle_geom* g, float** layers, int nlayers, double x, double y);

/* GridMapRosConverter::toOccupancyGrid (grid_map_ros/src/GridMapRosConverter.cpp:251-287).
 * out = int8[rows*cols] in nav_msgs/OccupancyGrid order. */
void oracle_to_occupancy(const oracle_geom* g, const float* layer, float data_min, float data_max,
                         signed char* out);

/* MapGlobalPlanner::ifBlocked (move_control/include/move_control/map_global_planner.h:39-54) over
 * grid_map::CircleIterator (grid_map_core/src/iterators/CircleIterator.cpp:17-37,80-113): 1 if any cell whose centre
 * lies within `radius` of (x, y) holds a non-NaN value > 0. */
int oracle_if_blocked(const oracle_geom* g, const float* master, double x, double y, double radius);

/* Steerer::update goal geometry (steerer.cpp:232-256): desiredDist (mm), desiredAngle (deg). */
void oracle_goal_from_pose(double rx, double ry, double yaw, double tx, double ty,
                           float* desired_angle, float* desired_dist);

/* The commented-out two-layer compose of MapProvider::composeMasterMapFromLayerdMap (map_provider.cpp:218-220). */
void oracle_compose_master(const float* range, const float* laser, float* master, long long n);

/* LaserScan -> RangeSamples: restatement of the specification in ros_navigation_b200/csrc/scan_project.h (intake side
 * of LaserMapUpdater::bufferIncomingMsg, laser_map_updater.cpp:38-144; laser_geometry / tf are not in the reference
 * tree: parity with the reference unpinned here). */
void oracle_sincos(double x, double* s, double* c);
/* simplifyLaserScan (laser_map_updater.cpp:114-144): sel[] = projected range indices (cap n_ranges + 1); returns n. */
int oracle_scan_select(float angle_increment, int n_ranges, int decimate, int* sel, float* increment_used);
/* Samples of one scan taken at sensor pose (x0, y0, yaw); dropped readings are skipped.  Returns the sample count. */
int oracle_project_scan(float angle_min, float increment_used, float range_min, float range_max, const int* sel,
                        int n_used, int decimated, const float* ranges, double x0, double y0, double yaw,
                        oracle_sample* out);

#ifdef __cplusplus
}
#endif
#endif

/*
 * vfh_tables.h -- host-side construction of the VFH+ lookup tables (the product's VFH::Init).
 *
 * Behavioural spec: move_control/src/vfh.cpp:53-110 (constructor), :130-231 (speed-dependent helpers),
 * :144-166 (SetCurrentMaxSpeed), :237-416 (Init).  The tables are flattened for the device:
 * only the rows in front of the robot (y < ceil(W/2)) are kept, cell-major f = y*W + x, and the per-cell
 * sector list becomes a bitmask of hist_size bits per (table, cell).
 *
 * Built on the host, like the reference, with the host libm (atanf/asinf/pow/tan): transcendental results
 * must come from the same libm as the reference build to be bit-identical.
 */
#ifndef B200NAV_VFH_TABLES_H
#define B200NAV_VFH_TABLES_H

#include <stdint.h>

#include <vector>

#include "../../include/b200nav.h"

namespace b200nav {

/* Scalars the kernel needs, with the reference's member types (vfh.h:300-340). */
struct VfhConst {
  float robot_radius;
  float cell_width;
  int window, center, front_rows; /* WINDOW_DIAMETER, CENTER_X(=CENTER_Y), ceil(W/2) */
  int hist_size, sector_angle, nwords;
  int num_tables;
  float safety_dist_0ms, safety_dist_1ms;
  int current_max_speed, max_speed;
  int max_speed_narrow, max_speed_wide;
  int max_acceleration;
  int max_turnrate_0ms, max_turnrate_1ms;
  float bin_low_0ms, bin_high_0ms, bin_low_1ms, bin_high_1ms;
  float u1, u2;
  double submap_length;
  float occupied_threshold;
};

struct VfhTables {
  VfhConst c;
  /* full W*W tables in the reference's [x][y] order (parity checks, b200nav_vfh_get_tables) */
  std::vector<float> dir_xy, dist_xy, base_xy;
  std::vector<uint32_t> masks_xy; /* [table][x][y][nwords] */
  /* device layout: front cells only, f = y*W + x */
  std::vector<float> dir, dist, base;
  std::vector<double> thr;        /* Cell_Dist + CELL_WIDTH/2.0 (vfh.cpp:1017)                    */
  std::vector<int16_t> kidx;      /* (int)rint(Cell_Direction*2.0) (vfh.cpp:1018); -2 at centre   */
  std::vector<uint32_t> masks;    /* [table][f][nwords]                                           */
  std::vector<int32_t> min_turning_radius; /* [0..current_max_speed] (vfh.cpp:144-166)            */
};

/* Validates p; returns 0 or a B200NAV_E* code with a message in err. */
int vfh_build_tables(const b200nav_vfh_params& p, VfhTables& t, char* err, int errlen);
void vfh_build_min_turning_radius(VfhTables& t, const b200nav_vfh_params& p, int max_speed);

/* Speed-dependent helpers shared with the kernel (same expressions on both sides). */
int vfh_get_max_turnrate(int max_turnrate_0ms, int max_turnrate_1ms, int speed);

}  // namespace b200nav
#endif

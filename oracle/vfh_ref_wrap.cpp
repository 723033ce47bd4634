/*
 * vfh_ref_wrap.cpp -- C wrapper around the UNMODIFIED reference VFH+ class (test infrastructure).
 *
 * Linked together with /root/reference/move_control/src/vfh.cpp (compiled where it lies, never copied)
 * into oracle/_ref/libvfh_ref.so by oracle/Makefile.  The reference class is the VFH+ oracle:
 *   - move_control::VFH            move_control/include/move_control/vfh.h:182-361
 *   - VFH::Init / Update_VFH       move_control/src/vfh.cpp:237-416, 480-605
 * The wrapper only (a) makes the wall clock deterministic (--wrap=gettimeofday; Update_VFH derives its
 * acceleration step from gettimeofday, vfh.cpp:521-531), (b) exposes private state for comparison, and
 * (c) pins the three reads of indeterminate memory the reference performs (SURVEY H4 a, e):
 *   - odd WINDOW_DIAMETER reads laser_ranges[-2] for the centre cell (vfh.cpp:1018 with Cell_Direction=-1):
 *     the wrapper hands the class a buffer whose entry [-2] is +inf, so the centre cell contributes 0;
 *   - Blocked_Circle_Radius is not initialised by the constructor (vfh.cpp:72-95) but read by
 *     Cant_Turn_To_Goal (vfh.cpp:638-647): set to 0 after construction;
 *   - Max_Speed_For_Picked_Angle is not initialised either: set to 0.
 */
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <sys/time.h>
#include <vector>

#define private public
#include "move_control/vfh.h"
#undef private

static double g_fake_time = 1000.0;

extern "C" int __wrap_gettimeofday(struct timeval* tv, void* /*tz*/) {
  const double t = g_fake_time;
  tv->tv_sec = static_cast<time_t>(std::floor(t));
  tv->tv_usec = static_cast<suseconds_t>(std::llround((t - std::floor(t)) * 1e6));
  if (tv->tv_usec >= 1000000) {
    tv->tv_sec += 1;
    tv->tv_usec -= 1000000;
  }
  return 0;
}

using move_control::VFH;

extern "C" {

/* params: the 19 constructor arguments in order (vfh.h:185-203), then robot_radius. */
void* vfhref_create(const double* p) {
  VFH* v = new VFH(p[0], (int)p[1], (int)p[2], p[3], p[4], (int)p[5], (int)p[6], (int)p[7], (int)p[8], (int)p[9],
                   (int)p[10], (int)p[11], p[12], p[13], p[14], p[15], p[16], p[17], p[18]);
  v->SetRobotRadius((float)p[19]);
  v->Blocked_Circle_Radius = 0.0f;
  v->Max_Speed_For_Picked_Angle = 0;
  v->Init();
  return v;
}

void vfhref_destroy(void* h) { delete static_cast<VFH*>(h); }

/* Absolute fake wall-clock time in seconds (microsecond resolution, like gettimeofday). */
void vfhref_set_time(double t) { g_fake_time = t; }
double vfhref_get_time(void) { return g_fake_time; }

void vfhref_set_current_max_speed(void* h, int s) { static_cast<VFH*>(h)->SetCurrentMaxSpeed(s); }

/* ranges: double[361][2]. */
int vfhref_update(void* h, const double* ranges, int current_speed, float goal_dir, float goal_dist, float tol,
                  int* chosen_speed, int* chosen_turnrate) {
  double buf[363][2];
  buf[0][0] = buf[0][1] = std::numeric_limits<double>::infinity();
  buf[1][0] = buf[1][1] = std::numeric_limits<double>::infinity();
  std::memcpy(&buf[2][0], ranges, sizeof(double) * 361 * 2);
  int cs = 0, ct = 0;
  const int rc = static_cast<VFH*>(h)->Update_VFH(&buf[2], current_speed, goal_dir, goal_dist, tol, cs, ct);
  *chosen_speed = cs;
  *chosen_turnrate = ct;
  return rc;
}

int vfhref_hist_size(void* h) { return static_cast<VFH*>(h)->HIST_SIZE; }
int vfhref_window(void* h) { return static_cast<VFH*>(h)->WINDOW_DIAMETER; }
int vfhref_num_tables(void* h) { return static_cast<VFH*>(h)->NUM_CELL_SECTOR_TABLES; }

/* State after the last update.  f = {Picked_Angle, Last_Picked_Angle, Desired_Angle, Blocked_Circle_Radius},
 * i = {last_chosen_speed, Max_Speed_For_Picked_Angle, Current_Max_Speed}. */
void vfhref_get_state(void* h, float* origin_hist, float* hist, float* last_binary, float* f, int* i) {
  VFH* v = static_cast<VFH*>(h);
  const int n = v->HIST_SIZE;
  if (origin_hist) std::memcpy(origin_hist, v->OriginHist, sizeof(float) * n);
  if (hist) std::memcpy(hist, v->Hist, sizeof(float) * n);
  if (last_binary) std::memcpy(last_binary, v->Last_Binary_Hist, sizeof(float) * n);
  if (f) {
    f[0] = v->Picked_Angle;
    f[1] = v->Last_Picked_Angle;
    f[2] = v->Desired_Angle;
    f[3] = v->Blocked_Circle_Radius;
  }
  if (i) {
    i[0] = v->last_chosen_speed;
    i[1] = v->Max_Speed_For_Picked_Angle;
    i[2] = v->Current_Max_Speed;
  }
}

/* Tables built by Init (vfh.cpp:237-416), row-major [x][y] as W*W floats. */
void vfhref_get_cell_tables(void* h, float* dir, float* dist, float* base_mag) {
  VFH* v = static_cast<VFH*>(h);
  const int W = v->WINDOW_DIAMETER;
  for (int x = 0; x < W; x++)
    for (int y = 0; y < W; y++) {
      if (dir) dir[x * W + y] = v->Cell_Direction[x][y];
      if (dist) dist[x * W + y] = v->Cell_Dist[x][y];
      if (base_mag) base_mag[x * W + y] = v->Cell_Base_Mag[x][y];
    }
}

/* Cell_Sector[table][x][y] as a bitmask of nwords 32-bit words per cell; returns 0 if any list is not strictly
 * ascending (then the mask would lose ordering information), else 1. */
int vfhref_get_sector_masks(void* h, int table, unsigned* masks, int nwords) {
  VFH* v = static_cast<VFH*>(h);
  const int W = v->WINDOW_DIAMETER;
  int ok = 1;
  for (int x = 0; x < W; x++)
    for (int y = 0; y < W; y++) {
      unsigned* m = masks + (size_t)(x * W + y) * nwords;
      for (int k = 0; k < nwords; k++) m[k] = 0;
      int prev = -1;
      for (int s : v->Cell_Sector[table][x][y]) {
        if (s <= prev) ok = 0;
        prev = s;
        m[s >> 5] |= 1u << (s & 31);
      }
    }
  return ok;
}

void vfhref_get_min_turning_radius(void* h, int* out, int n) {
  VFH* v = static_cast<VFH*>(h);
  for (int i = 0; i < n && i < (int)v->Min_Turning_Radius.size(); i++) out[i] = v->Min_Turning_Radius[i];
}

void vfhref_get_cell_mag(void* h, float* mag) {
  VFH* v = static_cast<VFH*>(h);
  const int W = v->WINDOW_DIAMETER;
  for (int x = 0; x < W; x++)
    for (int y = 0; y < W; y++) mag[x * W + y] = v->Cell_Mag[x][y];
}

} /* extern "C" */

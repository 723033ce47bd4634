/*
 * steer_host.cpp -- the caller glue of Steerer::update (move_control/src/steerer.cpp:221-258) for fleets: waypoint
 * advance and the goal geometry that VFH::Update_VFH takes, for n robots in one call.
 *
 * This part of the reference is host code around tf / odometry look-ups and stays host code here: it is a handful of
 * float operations per robot whose results depend on the platform's libm (hypotf, atan2f - glibc's atan2f is not
 * correctly rounded, so no device restatement could promise the reference's bits), and computing it with the same
 * libm calls as the reference is what keeps goal_direction identical.  It writes straight into the b200nav_vfh_input
 * records that b200nav_vfh_update_batched[_async] consumes (pinned memory: no extra copy).
 */
#include <math.h>
#include <stdint.h>

#include "../../include/b200nav.h"

#define B200NAV_RAD2DEG(a) ((a) * 180.0 / M_PI) /* steerer.cpp:12 */

namespace {
/* angles::normalize_angle_positive (header-only upstream package `angles`) */
inline double normalize_angle_positive(double a) { return fmod(fmod(a, 2.0 * M_PI) + 2.0 * M_PI, 2.0 * M_PI); }
}  // namespace

extern "C" int b200nav_steer_update_goals(int n, const double* poses, const double* waypoints,
                                          const int32_t* wp_offsets, int32_t* plan_index, float tolerance,
                                          const double* odom_speed, b200nav_vfh_input* inputs, uint8_t* plan_done) {
  if (n < 0 || !poses || !waypoints || !wp_offsets || !plan_index || !inputs) return B200NAV_EINVAL;
  for (int r = 0; r < n; r++) {
    const double cx = poses[3 * r], cy = poses[3 * r + 1], dir = poses[3 * r + 2];
    const int first = wp_offsets[r], size = wp_offsets[r + 1] - wp_offsets[r];
    b200nav_vfh_input& in = inputs[r];
    in.x = cx;
    in.y = cy;
    in.yaw = dir;
    in.goal_tolerance = tolerance;
    /* (int)(currentOdom.twist.twist.linear.x * 1000.0), steerer.cpp:252,262 */
    in.current_speed = odom_speed ? (int)(odom_speed[r] * 1000.0) : 0;
    bool done = plan_index[r] >= size || plan_index[r] < 0;
    float dx = 0.f, dy = 0.f, dist = 0.f;
    while (!done) { /* steerer.cpp:234-250: skip the waypoints that are already within the tolerance */
      const double* target = waypoints + 2 * (size_t)(first + plan_index[r]);
      dx = (target[0] - cx) * 1000.0;
      dy = (target[1] - cy) * 1000.0;
      dist = hypot(dx, dy); /* float overload, as in the reference (C++ <math.h>) */
      if (dist < tolerance) {
        plan_index[r]++;
        if (plan_index[r] >= size) done = true; /* ifPlanReady_ = false; return */
      } else {
        break;
      }
    }
    if (plan_done) plan_done[r] = done ? 1 : 0;
    if (done) { /* the reference publishes nothing for this robot in this cycle: goal = straight ahead, at the goal */
      in.goal_direction = 90.f;
      in.goal_distance = 0.f;
      continue;
    }
    in.goal_distance = dist;
    in.goal_direction = B200NAV_RAD2DEG(normalize_angle_positive(atan2(dy, dx) - dir + M_PI / 2)); /* steerer.cpp:254 */
  }
  return B200NAV_OK;
}

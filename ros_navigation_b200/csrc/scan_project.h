/*
 * scan_project.h -- LaserScan -> RangeSample on the device.
 *
 * Replaces (behaviour, not code) the intake side of LaserMapUpdater::bufferIncomingMsg
 * (move_control/src/laser_map_updater.cpp:38-144): getLaserOriginOnGlobal, simplifyLaserScan,
 * laser_geometry::LaserProjection::transformLaserScanToPointCloud and the per-point loop that fills RangeSamples.
 * laser_geometry and tf are third-party code that is NOT in the reference tree (ROS indigo, un-vendored), so their
 * arithmetic is restated here as an explicit specification - parity with the reference is UNPINNED at this boundary
 * (SURVEY section 8c); parity with the test suite's independent CPU restatement of this same specification is bit-exact:
 *
 *   selection   simplifyLaserScan (laser_map_updater.cpp:126-144): if angle_increment < 0.017 keep ranges[0] and then
 *               every ranges[i] at which the float accumulator `increment += angle_increment` reaches 0.017 (reset to
 *               0); the projected scan uses the last accumulated value as ITS angle_increment.  b200nav_scan_select.
 *   local point for the j-th selected range r (float32): dropped unless range_min <= r < range_max (laser_geometry
 *               projectLaser); angle = (double)angle_min + (double)j * (double)increment_used;
 *               lx = (float)((double)r * cos(angle)), ly = (float)((double)r * sin(angle))   (PointCloud2 floats)
 *   map frame   sensor pose (x0, y0, yaw) in the map frame at the scan's stamp (one transform per scan: the per-point
 *               time interpolation of the high-fidelity projection is not modelled):
 *               X = (float)((cos(yaw) * (double)lx - sin(yaw) * (double)ly) + x0), Y likewise with (sin, cos) and y0.
 *   sample      start = (x0, y0) (getLaserOriginOnGlobal), end = ((double)X, (double)Y).  ifClearEnd
 *               (laser_map_updater.cpp:62-66) is read from msg->ranges[index] with `index` the point's position in the
 *               PROJECTED scan but `msg` the ORIGINAL one: for a scan that was not thinned that reading passed the
 *               range filter, so the flag is false; for a thinned scan the j-th selected reading gets
 *               ifClearEnd = isinf(ranges[j]) || ranges[j] == range_max of the original scan (a reference quirk, kept:
 *               verified against the reference's own bufferIncomingMsg, tests/test_reference_pin.py).
 *   sin / cos   b200nav_sincos below: a fixed sequence of individually rounded fp64 operations (Cody-Waite reduction
 *               by pi/2 in three 33-bit pieces, the classic degree-13 / degree-14 minimax kernels), so that a CPU
 *               restatement and the GPU produce the same bits.  |x| < 2^20 * pi/2; about 1 ulp.
 */
#ifndef B200NAV_SCAN_PROJECT_H
#define B200NAV_SCAN_PROJECT_H

#include "geometry.h"

namespace b200nav {

B200_HD void b200nav_sincos(double x, double& s, double& c) {
  const double inv_pio2 = 6.36619772367581382433e-01;
  const double p1 = 1.57079632673412561417e+00; /* first 33 bits of pi/2 */
  const double p2 = 6.07710050630396597660e-11; /* next 33 bits          */
  const double p3 = 2.02226624871116645580e-21; /* next 33 bits          */
  const double magic = 6755399441055744.0;      /* 1.5 * 2^52: (t + magic) - magic rounds t to the nearest integer */
  const double fn = (x * inv_pio2 + magic) - magic;
  const double r = ((x - fn * p1) - fn * p2) - fn * p3;
  const double z = r * r;
  const double ps = -1.66666666666666324348e-01 +
                    z * (8.33333333332248946124e-03 +
                         z * (-1.98412698298579493134e-04 +
                              z * (2.75573137070700676789e-06 +
                                   z * (-2.50507602534068634195e-08 + z * 1.58969099521155010221e-10))));
  const double sk = r + r * (z * ps);
  const double pc = 4.16666666666666019037e-02 +
                    z * (-1.38888888888741095749e-03 +
                         z * (2.48015872894767294178e-05 +
                              z * (-2.75573143513906633035e-07 +
                                   z * (2.08757232129817482790e-09 + z * -1.13596475577881948265e-11))));
  const double ck = 1.0 - (0.5 * z - (z * z) * pc);
  const long long n = (long long)fn;
  const int q = (int)(n & 3);
  s = (q == 0) ? sk : (q == 1) ? ck : (q == 2) ? -sk : -ck;
  c = (q == 0) ? ck : (q == 1) ? -sk : (q == 2) ? -ck : sk;
}

/* Scan geometry shared by all robots of an update (sensor_msgs/LaserScan header fields are float32). */
struct ScanModel {
  float angle_min;
  float increment_used; /* angle_increment of the (possibly simplified) scan */
  float range_min, range_max;
  int n_ranges;         /* ranges per robot in the input array                */
  int n_used;           /* selected ranges per robot (== samples per robot)   */
  int decimated;        /* simplifyLaserScan thinned the scan (angle_increment < 0.017) */
};

/* ifClearEnd of the j-th projected point: laser_map_updater.cpp:62-66 indexes the ORIGINAL scan with j. */
B200_HD bool scan_clear_end(const ScanModel& m, float original_range_at_j) {
  return m.decimated && (original_range_at_j == m.range_max || original_range_at_j == INFINITY ||
                         original_range_at_j == -INFINITY);
}

/* j-th selected range of one scan -> sample end point; returns false when the reading is dropped. */
B200_HD bool project_reading(const ScanModel& m, int j, float r, double x0, double y0, double cy, double sy,
                             double& ex, double& ey) {
  if (!(r >= m.range_min && r < m.range_max)) return false; /* also drops NaN and inf */
  const double angle = (double)m.angle_min + (double)j * (double)m.increment_used;
  double sa, ca;
  b200nav_sincos(angle, sa, ca);
  const float lx = (float)((double)r * ca), ly = (float)((double)r * sa);
  const float X = (float)((cy * (double)lx - sy * (double)ly) + x0);
  const float Y = (float)((sy * (double)lx + cy * (double)ly) + y0);
  ex = (double)X;
  ey = (double)Y;
  return true;
}

}  // namespace b200nav
#endif

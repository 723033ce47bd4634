/*
 * shim_check.cpp -- drives the C++ drop-in classes of include/b200nav_shim.hpp the way the reference node does
 * (nav_only_vfh_node: MapProvider::updateMap at 5 Hz -> Steerer::update) and compares every cycle with the CPU oracle:
 *   layer "laser"/"master" vs oracle_himm_update                      (bit-exact)
 *   Update_VFH_FromGrid vs oracle_ranges_from_submap + reference VFH  (commands, Hist, OriginHist exact)
 *   Update_VFH (host pseudo-scan, the unmodified reference call)      (same)
 * Needs a GPU.  Prints "OK ..." or the first mismatch.
 */
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/b200nav_shim.hpp"
#include "../../oracle/oracle_api.h"

extern "C" {
void* vfhref_create(const double* p);
void vfhref_destroy(void*);
void vfhref_set_time(double t);
int vfhref_update(void*, const double*, int, float, float, float, int*, int*);
void vfhref_get_state(void*, float*, float*, float*, float*, int*);
}

static bool same_layer(const std::vector<float>& a, const std::vector<float>& b) {
  for (size_t i = 0; i < a.size(); i++) {
    const bool na = std::isnan(a[i]), nb = std::isnan(b[i]);
    if (na != nb || (!na && std::memcmp(&a[i], &b[i], 4) != 0)) {
      printf("MISMATCH: layer cell %zu: %g vs %g\n", i, a[i], b[i]);
      return false;
    }
  }
  return true;
}

/* range of a ray from (x,y) at angle th inside a square room [-4,4]^2 with a disc obstacle */
static double cast(double x, double y, double th) {
  const double dx = std::cos(th), dy = std::sin(th), R = 4.0;
  double t = 1e9;
  if (dx > 0) t = std::min(t, (R - x) / dx);
  if (dx < 0) t = std::min(t, (-R - x) / dx);
  if (dy > 0) t = std::min(t, (R - y) / dy);
  if (dy < 0) t = std::min(t, (-R - y) / dy);
  const double cx = 1.2, cy = 0.6, r = 0.35, fx = x - cx, fy = y - cy;
  const double b = fx * dx + fy * dy, c = fx * fx + fy * fy - r * r, disc = b * b - c;
  if (disc > 0) {
    const double td = -b - std::sqrt(disc);
    if (td > 0) t = std::min(t, td);
  }
  return t;
}

int main() {
  const double L = 10.0, res = 0.05;
  b200nav::Context ctx(0);
  b200nav::GridLayers map(ctx, L, L, res);
  map.add("master");
  b200nav::LaserMapUpdater laser(map, "laser");
  /* Steerer::initVfh defaults (steerer.cpp:69-121) */
  b200nav::VFH vfh(ctx, 100, 30, 5, 10, 50, 200, 200, 300, 200, 40, 40, 40, 1.0, 2e6, 4e6, 2e6, 4e6, 10.0, 1.0);
  vfh.SetRobotRadius(178.0f);
  if (!vfh.Init()) {
    printf("MISMATCH: Init failed: %s\n", ctx.last_error());
    return 1;
  }
  b200nav::VFH vfh2(ctx, 100, 30, 5, 10, 50, 200, 200, 300, 200, 40, 40, 40, 1.0, 2e6, 4e6, 2e6, 4e6, 10.0, 1.0);
  vfh2.SetRobotRadius(178.0f);
  vfh2.Init();

  oracle_geom g;
  oracle_geom_init(&g, L, L, res, 0.0, 0.0);
  std::vector<float> o_laser((size_t)g.rows * g.cols, NAN), o_master(o_laser), d_layer(o_laser.size());
  const double params[20] = {100, 30, 5, 10, 50, 200, 200, 300, 200, 40, 40, 40, 1.0, 2e6, 4e6, 2e6, 4e6, 10.0, 1.0, 178.0};
  vfhref_set_time(1000.0);
  void* ref = vfhref_create(params);
  void* ref2 = vfhref_create(params);
  double now = 1000.0;

  int speed = 0, speed2 = 0;
  for (int step = 0; step < 120; step++) {
    const double t = step * 0.2;
    const double x = 2.0 * std::sin(0.05 * t), y = 1.5 * std::sin(0.08 * t + 1.0), yaw = 0.3 * t;
    /* scan intake (LaserMapUpdater::bufferIncomingMsg): 360 beams, range_max 3 m, float32 cloud points */
    std::vector<oracle_sample> samples;
    for (int i = 0; i < 360; i++) {
      const double th = yaw + (-M_PI + 2 * M_PI * i / 360.0);
      const double r = cast(x, y, th);
      if (r >= 3.0) continue; /* laser_geometry drops range_max readings */
      oracle_sample s;
      s.sx = x;
      s.sy = y;
      s.ex = (double)(float)(x + r * std::cos(th));
      s.ey = (double)(float)(y + r * std::sin(th));
      s.clear_end = 0;
      s.pad_ = 0;
      samples.push_back(s);
      laser.pushSample(s.sx, s.sy, s.ex, s.ey, false);
    }
    /* MapProvider::updateMap (map_provider.cpp:190-205) */
    double minX = 0, minY = 0, maxX = 0, maxY = 0, bb[4] = {0, 0, 0, 0};
    laser.updateMap(minX, minY, maxX, maxY);
    map.copyLayer("master", "laser");
    oracle_himm_update(&g, o_laser.data(), samples.data(), (int)samples.size(), bb);
    o_master = o_laser;
    if (bb[0] != minX || bb[1] != minY || bb[2] != maxX || bb[3] != maxY) {
      printf("MISMATCH: bbox at step %d\n", step);
      return 1;
    }
    /* Steerer::update (steerer.cpp:221-270) */
    float gdir, gdist;
    oracle_goal_from_pose(x, y, yaw, 3.0, 2.0, &gdir, &gdist);
    double ranges[361][2];
    oracle_ranges_from_submap(&g, o_master.data(), x, y, yaw, 1.5, &ranges[0][0]);
    now += 0.2;
    vfhref_set_time(now);
    int rs, rt, rs2, rt2;
    vfhref_update(ref, &ranges[0][0], speed, gdir, gdist, 250.f, &rs, &rt);
    vfhref_update(ref2, &ranges[0][0], speed2, gdir, gdist, 250.f, &rs2, &rt2);
    int cs = -1, ct = -1, cs2 = -1, ct2 = -1;
    vfh.SetNextElapsed(0.2);
    vfh.Update_VFH_FromGrid(map, "master", x, y, yaw, speed, gdir, gdist, 250.f, cs, ct);
    vfh2.SetNextElapsed(0.2);
    vfh2.Update_VFH(ranges, speed2, gdir, gdist, 250.f, cs2, ct2);
    if (cs != rs || ct != rt || cs2 != rs2 || ct2 != rt2) {
      printf("MISMATCH: command at step %d: grid (%d,%d) ranges (%d,%d) reference (%d,%d)\n", step, cs, ct, cs2, ct2, rs,
             rt);
      return 1;
    }
    float oh[72], h[72], fl[4];
    vfhref_get_state(ref, oh, h, nullptr, fl, nullptr);
    if (std::memcmp(oh, vfh.OriginHist, sizeof(oh)) || std::memcmp(h, vfh.Hist, sizeof(h)) ||
        std::memcmp(oh, vfh2.OriginHist, sizeof(oh)) || fl[0] != vfh.GetPickedAngle()) {
      printf("MISMATCH: histogram / picked angle at step %d\n", step);
      return 1;
    }
    speed = rs;
    speed2 = rs2;
  }
  map.download("master", d_layer.data());
  if (!same_layer(d_layer, o_master)) return 1;
  /* toOccupancyGrid */
  std::vector<int8_t> occ_d(o_master.size()), occ_o(o_master.size());
  map.toOccupancyGrid("master", 0.0f, 255.0f, occ_d.data());
  oracle_to_occupancy(&g, o_master.data(), 0.0f, 255.0f, (signed char*)occ_o.data());
  if (std::memcmp(occ_d.data(), occ_o.data(), occ_o.size())) {
    printf("MISMATCH: occupancy grid\n");
    return 1;
  }
  vfhref_destroy(ref);
  vfhref_destroy(ref2);
  printf("OK shim: 120 cycles, layers bit-exact, commands and histograms equal to the reference\n");
  return 0;
}

// nav_harness.cpp -- C driver around the reference's OWN node classes (TEST INFRASTRUCTURE; see oracle/oracle_api.h).
//
// Compiled twice from this one file, against the stand-in middleware headers of oracle/standin/:
//
//   oracle/_ref/libnav_ref.so          with the reference's unmodified sources compiled where they lie
//                                      (move_control/src/{map_provider,steerer,laser_map_updater,range_map_updater,
//                                      vfh}.cpp, grid_map_core/src/*.cpp): the reference itself, runnable here.
//   tests/cpp/_build/libnav_dropin.so  with the SAME map_provider.cpp / steerer.cpp / grid_map_core, but the updater
//                                      and VFH classes taken from the product's drop-in headers
//                                      (include/move_control/*.h -> libb200nav.so).  -DNAVH_DROPIN.
//
// Both export the same navh_* functions, so one test script drives the two builds through identical scenarios and
// compares grids, pseudo-scans, histograms and velocity commands.  The harness reaches private members
// (MapProvider::updateMap, map_, Steerer::update, ranges_, vfhP_) through the usual test trick below; it adds no
// behaviour of its own apart from calling the bodies of the reference's timer loops once per request
// (loopUpdateAndPublishMap: map_provider.cpp:151-175, loopMoveMap: :177-188, vfhLoop: steerer.cpp:135-144).
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <sys/time.h>

#include <algorithm>
#include <functional>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "standin_ros_core.hpp"
#include <Eigen/Core>

#define private public
#define protected public
#include "move_control/map_provider.h"
#include "move_control/map_updater.h"
#include "move_control/steerer.h"
#ifndef NAVH_DROPIN
#include "move_control/laser_map_updater.h"
#endif
#undef private
#undef protected

#define NAVH_API extern "C" __attribute__((visibility("default")))

// Update_VFH measures elapsed wall time (vfh.cpp:521-531); both builds are linked with -Wl,--wrap=gettimeofday so
// that the world's clock answers instead.
extern "C" int __real_gettimeofday(struct timeval* tv, void* tz);
extern "C" __attribute__((visibility("default"))) int __wrap_gettimeofday(struct timeval* tv, void* tz) {
  standin::World* w = standin::current();
  if (!w) return __real_gettimeofday(tv, tz);
  tv->tv_sec = (time_t)(w->now_ns / 1000000000LL) + 1000000;
  tv->tv_usec = (suseconds_t)((w->now_ns % 1000000000LL) / 1000);
  return 0;
}

namespace {

struct Nav {
  standin::World world;
  std::unique_ptr<ros::NodeHandle> nh;
  std::unique_ptr<move_control::MapProvider> provider;
  std::unique_ptr<move_control::Steerer> steerer;
};

void set_frame(Nav* n, const char* frame, double x, double y, double yaw) {
  standin::Frame f;
  f.x = x;
  f.y = y;
  f.yaw = yaw;
  n->world.frames[standin::World::norm(frame)] = f;
}

void publish_scan(Nav* n, const char* topic, const char* frame, float angle_min, float angle_increment,
                  float range_min, float range_max, const float* ranges, int count) {
  sensor_msgs::LaserScan s;
  s.header.stamp = ros::Time::now();
  s.header.frame_id = frame;
  s.angle_min = angle_min;
  s.angle_increment = angle_increment;
  s.angle_max = angle_min + angle_increment * (float)(count - 1);
  s.range_min = range_min;
  s.range_max = range_max;
  s.ranges.assign(ranges, ranges + count);
  n->world.publish<sensor_msgs::LaserScan>(topic, s);
}

}  // namespace

NAVH_API int navh_is_dropin() {
#ifdef NAVH_DROPIN
  return 1;
#else
  return 0;
#endif
}

// MapProvider(nh, Length(len_x, len_y), moving) + Steerer(nh, provider), as the reference's nodes construct them
// (nav_only_vfh_node.cpp:40-41).  params: ROS parameters Steerer::initVfh reads (steerer.cpp:73-121).
NAVH_API void* navh_create(double len_x, double len_y, int moving, const char** keys, const double* vals, int nparams,
                           double t0) {
  Nav* n = new Nav();
  n->world.now_ns = ros::Time(t0).ns;
  for (int i = 0; i < nparams; ++i) n->world.params[keys[i]] = vals[i];
  set_frame(n, "odom", 0, 0, 0);
  set_frame(n, "base_link", 0, 0, 0);
  standin::Scope scope(&n->world);
  try {
    n->nh.reset(new ros::NodeHandle());
    n->provider.reset(new move_control::MapProvider(*n->nh, grid_map::Length(len_x, len_y), moving != 0));
    n->steerer.reset(new move_control::Steerer(*n->nh, *n->provider));
  } catch (const std::exception& e) {
    fprintf(stderr, "navh_create: %s\n", e.what());
    delete n;
    return nullptr;
  }
  return n;
}

NAVH_API void navh_destroy(void* h) {
  Nav* n = (Nav*)h;
  if (!n) return;
  standin::Scope scope(&n->world);
  n->steerer.reset();
  n->provider.reset();
  n->nh.reset();
  delete n;
}

NAVH_API void navh_set_time(void* h, double t) { ((Nav*)h)->world.now_ns = ros::Time(t).ns; }
NAVH_API void navh_set_frame(void* h, const char* frame, double x, double y, double yaw) {
  set_frame((Nav*)h, frame, x, y, yaw);
}

NAVH_API void navh_publish_scan(void* h, const char* topic, const char* frame, float angle_min, float angle_increment,
                                float range_min, float range_max, const float* ranges, int count) {
  Nav* n = (Nav*)h;
  standin::Scope scope(&n->world);
  publish_scan(n, topic, frame, angle_min, angle_increment, range_min, range_max, ranges, count);
}

NAVH_API void navh_publish_range(void* h, const char* topic, const char* frame, float range, float min_range,
                                 float max_range) {
  Nav* n = (Nav*)h;
  standin::Scope scope(&n->world);
  sensor_msgs::Range r;
  r.header.stamp = ros::Time::now();
  r.header.frame_id = frame;
  r.range = range;
  r.min_range = min_range;
  r.max_range = max_range;
  n->world.publish<sensor_msgs::Range>(topic, r);
}

NAVH_API void navh_publish_odom(void* h, double linear_x) {
  Nav* n = (Nav*)h;
  standin::Scope scope(&n->world);
  nav_msgs::Odometry o;
  o.header.stamp = ros::Time::now();
  o.twist.twist.linear.x = linear_x;
  n->world.publish<nav_msgs::Odometry>("/odom", o);
}

// one pass of loopUpdateAndPublishMap's body without the publish (map_provider.cpp:158)
NAVH_API void navh_update_map(void* h) {
  Nav* n = (Nav*)h;
  standin::Scope scope(&n->world);
  n->provider->updateMap();
}

// one pass of loopMoveMap's body (map_provider.cpp:183-185)
NAVH_API int navh_move_map(void* h) {
  Nav* n = (Nav*)h;
  standin::Scope scope(&n->world);
  grid_map::Position p;
  if (!n->provider->getRobotPos(p)) return -1;
  return n->provider->map_.move(p) ? 1 : 0;
}

// MapProvider::publishMap() (map_provider.cpp:207-214) -> the OccupancyGrid that reached "global_map"
NAVH_API int navh_publish_map(void* h, int8_t* out, int cap) {
  Nav* n = (Nav*)h;
  standin::Scope scope(&n->world);
  n->provider->publishMap();
  auto m = n->world.last_of<nav_msgs::OccupancyGrid>("global_map");
  if (!m) return -1;
  const int cnt = (int)m->data.size();
  if (out && cap >= cnt) memcpy(out, m->data.data(), cnt);
  return cnt;
}

// geo = {rows, cols, start0, start1}; dbl = {pos_x, pos_y, len_x, len_y, res}
NAVH_API void navh_geometry(void* h, int* geo, double* dbl) {
  Nav* n = (Nav*)h;
  const grid_map::GridMap& m = n->provider->map_;
  geo[0] = m.getSize()(0);
  geo[1] = m.getSize()(1);
  geo[2] = m.getStartIndex()(0);
  geo[3] = m.getStartIndex()(1);
  dbl[0] = m.getPosition()(0);
  dbl[1] = m.getPosition()(1);
  dbl[2] = m.getLength()(0);
  dbl[3] = m.getLength()(1);
  dbl[4] = m.getResolution();
}

// layer as the host GridMap holds it: rows*cols floats, column-major (Eigen::MatrixXf::data())
NAVH_API int navh_get_layer(void* h, const char* layer, float* out, int cap) {
  Nav* n = (Nav*)h;
  grid_map::GridMap& m = n->provider->map_;
  if (!m.exists(layer)) return -1;
  const grid_map::Matrix& d = m[layer];
  const int cnt = (int)(d.rows() * d.cols());
  if (out && cap >= cnt) memcpy(out, d.data(), sizeof(float) * cnt);
  return cnt;
}

NAVH_API void navh_accept_plan(void* h, const double* xy, int count) {
  Nav* n = (Nav*)h;
  standin::Scope scope(&n->world);
  std::vector<grid_map::Position> plan;
  for (int i = 0; i < count; ++i) plan.push_back(grid_map::Position(xy[2 * i], xy[2 * i + 1]));
  n->steerer->acceptPlan(plan);
}

NAVH_API int navh_robot_pose(void* h, double* xyyaw) {
  Nav* n = (Nav*)h;
  standin::Scope scope(&n->world);
  grid_map::Position p;
  double yaw = 0;
  if (!n->provider->getRobotPos(p, yaw)) return -1;
  xyyaw[0] = p(0);
  xyyaw[1] = p(1);
  xyyaw[2] = yaw;
  return 0;
}

typedef struct {
  double linear_x, angular_z; /* the Twist on /mobile_base/commands/velocity (Steerer::pubVel)                 */
  int32_t updated;            /* 1 if Steerer::update published a command in this call                       */
  int32_t plan_ready;         /* Steerer::ifPlanReady_ after the call                                        */
  float picked_angle;         /* VFH::GetPickedAngle                                                         */
  float desired_angle;        /* VFH::GetDesiredAngle                                                        */
  double ranges[361];         /* Steerer::ranges_[i][0]                                                      */
  float hist[72];             /* VFH::Hist (masked)                                                          */
  float origin_hist[72];      /* VFH::OriginHist (primary)                                                   */
} navh_steer_out;

// one pass of vfhLoop's body (steerer.cpp:140-141)
NAVH_API int navh_steer(void* h, navh_steer_out* out) {
  Nav* n = (Nav*)h;
  standin::Scope scope(&n->world);
  move_control::Steerer& s = *n->steerer;
  const long before = n->world.published["mobile_base/commands/velocity"];
  if (s.ifPlanReady_) s.update();
  memset(out, 0, sizeof(*out));
  out->updated = n->world.published["mobile_base/commands/velocity"] != before;
  out->plan_ready = s.ifPlanReady_ ? 1 : 0;
  auto tw = n->world.last_of<geometry_msgs::Twist>("/mobile_base/commands/velocity");
  if (tw) {
    out->linear_x = tw->linear.x;
    out->angular_z = tw->angular.z;
  }
  out->picked_angle = s.vfhP_->GetPickedAngle();
  out->desired_angle = s.vfhP_->GetDesiredAngle();
  for (int i = 0; i < 361; ++i) out->ranges[i] = s.ranges_[i][0];
  const int hs = std::min(72, s.vfhP_->getHistSize());
  for (int i = 0; i < hs; ++i) {
    out->hist[i] = s.vfhP_->Hist[i];
    out->origin_hist[i] = s.vfhP_->OriginHist[i];
  }
  return 0;
}

// the Histogram message of Steerer::pubHist (steerer.cpp:201-220): 36 front bins, uint16 (large values wrap)
NAVH_API int navh_last_hist_msg(void* h, uint16_t* ydata36, uint16_t* ybin36) {
  Nav* n = (Nav*)h;
  auto m = n->world.last_of<move_control::Histogram>("hist");
  if (!m) return -1;
  for (size_t i = 0; i < m->yData.size() && i < 36; ++i) {
    ydata36[i] = m->yData[i];
    ybin36[i] = m->yBinData[i];
  }
  return (int)m->num_bin;
}

#ifdef NAVH_DROPIN
// The optional fast path of INTEGRATION.md section 3 (drop-in build only): Steerer::update's goal glue through
// b200nav_steer_update_goals and VFH::Update_VFH_FromGrid on the device twin of MapProvider's map - no host submap
// copy, no host atan2.  `plan` / `plan_index` play Steerer::plan_ / planIndex_.  Returns 0 when a command was computed.
NAVH_API int navh_steer_from_grid(void* h, const double* plan_xy, int plan_n, int* plan_index, double odom_speed,
                                  const char* layer, navh_steer_out* out) {
  Nav* n = (Nav*)h;
  standin::Scope scope(&n->world);
  move_control::Steerer& s = *n->steerer;
  grid_map::Position pos;
  double yaw = 0;
  if (!n->provider->getRobotPos(pos, yaw)) return -1;
  const double pose[3] = {pos(0), pos(1), yaw};
  const int32_t offs[2] = {0, plan_n};
  b200nav_vfh_input in;
  memset(&in, 0, sizeof(in));
  uint8_t done = 0;
  int32_t idx = *plan_index;
  if (b200nav_steer_update_goals(1, pose, plan_xy, offs, &idx, 250.0f, &odom_speed, &in, &done) != B200NAV_OK) return -2;
  *plan_index = idx;
  memset(out, 0, sizeof(*out));
  out->plan_ready = done ? 0 : 1;
  if (done) return 1;
  int speed = 0, turn = 0;
  s.vfhP_->Update_VFH_FromGrid(n->provider->map_, layer, in.x, in.y, in.yaw, in.current_speed, in.goal_direction,
                               in.goal_distance, in.goal_tolerance, speed, turn);
  out->updated = 1;
  out->linear_x = (float)(speed) / 1000.0;          /* Steerer::pubVel (steerer.cpp:193-199) */
  out->angular_z = turn * M_PI / 180.0;
  out->picked_angle = s.vfhP_->GetPickedAngle();
  out->desired_angle = s.vfhP_->GetDesiredAngle();
  const int hs = std::min(72, s.vfhP_->getHistSize());
  for (int i = 0; i < hs; ++i) {
    out->hist[i] = s.vfhP_->Hist[i];
    out->origin_hist[i] = s.vfhP_->OriginHist[i];
  }
  return 0;
}
#endif

// ---------------------------------------------------------------------------------------------------------------
// Fleet cycle: the reference node of every robot fed one scan and asked for one decision, robots block-partitioned
// over `threads` std::threads inside this one call (the CPU baseline of the batched configs).
//   poses: n * 3 doubles (sensor = robot pose in the map frame), ranges: n * count floats,
//   goals: n * 2 doubles (waypoint handed to Steerer::acceptPlan when no plan is active), speeds: n doubles (odom, m/s)
//   out_cmd: n * 2 doubles (linear.x, angular.z)
// ---------------------------------------------------------------------------------------------------------------
NAVH_API int navh_fleet_cycle(void** hs, int n, double t, const double* poses, const float* ranges, int count,
                              float angle_min, float angle_increment, float range_min, float range_max,
                              const double* goals, const double* speeds, double* out_cmd, int threads) {
  if (threads < 1) threads = 1;
  if (threads > n) threads = n;
  auto work = [&](int lo, int hi) {
    for (int r = lo; r < hi; ++r) {
      Nav* nav = (Nav*)hs[r];
      standin::Scope scope(&nav->world);
      nav->world.now_ns = ros::Time(t).ns;
      set_frame(nav, "base_link", poses[3 * r], poses[3 * r + 1], poses[3 * r + 2]);
      set_frame(nav, "laser", poses[3 * r], poses[3 * r + 1], poses[3 * r + 2]);
      publish_scan(nav, "/laser_scan", "laser", angle_min, angle_increment, range_min, range_max,
                   ranges + (size_t)r * count, count);
      nav->provider->updateMap();
      nav_msgs::Odometry o;
      o.twist.twist.linear.x = speeds ? speeds[r] : 0.0;
      nav->world.publish<nav_msgs::Odometry>("/odom", o);
      if (!nav->steerer->ifPlanReady_ && goals) {
        std::vector<grid_map::Position> plan;
        plan.push_back(grid_map::Position(poses[3 * r], poses[3 * r + 1]));
        plan.push_back(grid_map::Position(goals[2 * r], goals[2 * r + 1]));
        nav->steerer->acceptPlan(plan);
      }
      if (nav->steerer->ifPlanReady_) nav->steerer->update();
      auto tw = nav->world.last_of<geometry_msgs::Twist>("/mobile_base/commands/velocity");
      if (out_cmd) {
        out_cmd[2 * r] = tw ? tw->linear.x : 0.0;
        out_cmd[2 * r + 1] = tw ? tw->angular.z : 0.0;
      }
    }
  };
  if (threads == 1) {
    work(0, n);
    return 1;
  }
  std::vector<std::thread> pool;
  for (int k = 0; k < threads; ++k) {
    const int lo = (int)((long long)n * k / threads), hi = (int)((long long)n * (k + 1) / threads);
    pool.emplace_back(work, lo, hi);
  }
  for (auto& th : pool) th.join();
  return threads;
}

#ifndef NAVH_DROPIN
// Same cycle at the RangeSample boundary (the parity contract): the samples of every robot are put
// straight into its LaserMapUpdater's buffer (what bufferIncomingMsg would have pushed, laser_map_updater.cpp:57-70)
// - no thinning, every beam of the scan is applied, which is the workload the GPU arm runs - then
// MapProvider::updateMap (updaters + master compose) and Steerer::update (submap copy, pseudo-scan, VFH) run as written.
//   samples: 40-byte records {sx, sy, ex, ey (double), clear_end (int32), pad}; offsets: n + 1 ints
NAVH_API int navh_fleet_cycle_samples(void** hs, int n, double t, const double* poses, const void* samples,
                                      const int32_t* offsets, const double* goals, const double* speeds,
                                      double* out_cmd, int threads) {
  struct Rec {
    double sx, sy, ex, ey;
    int32_t clear_end, pad;
  };
  const Rec* recs = (const Rec*)samples;
  if (threads < 1) threads = 1;
  if (threads > n) threads = n;
  auto work = [&](int lo, int hi) {
    for (int r = lo; r < hi; ++r) {
      Nav* nav = (Nav*)hs[r];
      standin::Scope scope(&nav->world);
      nav->world.now_ns = ros::Time(t).ns;
      set_frame(nav, "base_link", poses[3 * r], poses[3 * r + 1], poses[3 * r + 2]);
      move_control::LaserMapUpdater* laser = nullptr;
      for (auto& u : nav->provider->mapUpdaters_)
        if (u->typeName_ == "laser") laser = static_cast<move_control::LaserMapUpdater*>(u.get());
      if (!laser) continue;
      {
        boost::unique_lock<boost::mutex> lock(laser->laserSampleBufferMutex_);
        for (int i = offsets[r]; i < offsets[r + 1]; ++i) {
          move_control::MapUpdater::RangeSample rs;
          rs.start = grid_map::Position(recs[i].sx, recs[i].sy);
          rs.end = grid_map::Position(recs[i].ex, recs[i].ey);
          rs.ifClearEnd = recs[i].clear_end != 0;
          laser->laserSampleBuffer_.push_back(rs);
        }
      }
      nav->provider->updateMap();
      nav_msgs::Odometry o;
      o.twist.twist.linear.x = speeds ? speeds[r] : 0.0;
      nav->world.publish<nav_msgs::Odometry>("/odom", o);
      if (goals) { /* a fresh two-point plan every cycle: the steering target of this cycle */
        std::vector<grid_map::Position> plan;
        plan.push_back(grid_map::Position(poses[3 * r], poses[3 * r + 1]));
        plan.push_back(grid_map::Position(goals[2 * r], goals[2 * r + 1]));
        nav->steerer->acceptPlan(plan);
      }
      if (nav->steerer->ifPlanReady_) nav->steerer->update();
      auto tw = nav->world.last_of<geometry_msgs::Twist>("/mobile_base/commands/velocity");
      if (out_cmd) {
        out_cmd[2 * r] = tw ? tw->linear.x : 0.0;
        out_cmd[2 * r + 1] = tw ? tw->angular.z : 0.0;
      }
    }
  };
  if (threads == 1) {
    work(0, n);
    return 1;
  }
  std::vector<std::thread> pool;
  for (int k = 0; k < threads; ++k) {
    const int lo = (int)((long long)n * k / threads), hi = (int)((long long)n * (k + 1) / threads);
    pool.emplace_back(work, lo, hi);
  }
  for (auto& th : pool) th.join();
  return threads;
}
#endif

NAVH_API int navh_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

// ---------------------------------------------------------------------------------------------------------------
// Core: the reference's MapUpdater::lineOnMap / clearCell / markCell (map_updater.h:38-71) on a grid_map::GridMap of
// ANY geometry (MapProvider hard-codes 5 cm cells) - what pins oracle_himm_update at the C2 / C3 geometries.
// ---------------------------------------------------------------------------------------------------------------
#ifndef NAVH_DROPIN
namespace {
struct CoreUpdater : public move_control::MapUpdater {
  CoreUpdater(ros::NodeHandle& nh, tf::TransformListener& tf, grid_map::GridMap& map, const std::string& name)
      : move_control::MapUpdater(nh, tf, map, name) {}
  void updateMap(double&, double&, double&, double&) override {}
  void addMonitorTopic(const std::string&) override {}
  void apply(const double* s5, int n, double* bbox) {
    for (int i = 0; i < n; ++i) {
      RangeSample rs;
      rs.start = grid_map::Position(s5[5 * i], s5[5 * i + 1]);
      rs.end = grid_map::Position(s5[5 * i + 2], s5[5 * i + 3]);
      rs.ifClearEnd = s5[5 * i + 4] != 0.0;
      lineOnMap(rs);
      if (bbox) { /* the two touch() calls of laser_map_updater.cpp:17-18 */
        touch(bbox[0], bbox[1], bbox[2], bbox[3], rs.start(0), rs.start(1));
        touch(bbox[0], bbox[1], bbox[2], bbox[3], rs.end(0), rs.end(1));
      }
    }
  }
};
struct Core {
  standin::World world;
  std::unique_ptr<ros::NodeHandle> nh;
  std::unique_ptr<tf::TransformListener> tf;
  grid_map::GridMap map;
  std::unique_ptr<CoreUpdater> up;
};
}  // namespace

NAVH_API void* navh_core_create(double len_x, double len_y, double res, double pos_x, double pos_y,
                                const char* layer) {
  Core* c = new Core();
  standin::Scope scope(&c->world);
  c->nh.reset(new ros::NodeHandle());
  c->tf.reset(new tf::TransformListener());
  c->map.setFrameId("odom");
  c->map.setGeometry(grid_map::Length(len_x, len_y), res, grid_map::Position(pos_x, pos_y));
  c->up.reset(new CoreUpdater(*c->nh, *c->tf, c->map, layer));
  return c;
}
NAVH_API void navh_core_destroy(void* h) { delete (Core*)h; }
NAVH_API void navh_core_size(void* h, int* rows, int* cols) {
  Core* c = (Core*)h;
  *rows = c->map.getSize()(0);
  *cols = c->map.getSize()(1);
}
// samples: n * 5 doubles {sx, sy, ex, ey, ifClearEnd}
NAVH_API void navh_core_update(void* h, const double* s5, int n, double* bbox) { ((Core*)h)->up->apply(s5, n, bbox); }
NAVH_API int navh_core_move(void* h, double x, double y) {
  return ((Core*)h)->map.move(grid_map::Position(x, y)) ? 1 : 0;
}
NAVH_API void navh_core_start_index(void* h, int* s2, double* pos2) {
  Core* c = (Core*)h;
  s2[0] = c->map.getStartIndex()(0);
  s2[1] = c->map.getStartIndex()(1);
  pos2[0] = c->map.getPosition()(0);
  pos2[1] = c->map.getPosition()(1);
}
NAVH_API int navh_core_get_layer(void* h, const char* layer, float* out, int cap) {
  Core* c = (Core*)h;
  if (!c->map.exists(layer)) return -1;
  const grid_map::Matrix& d = c->map[layer];
  const int cnt = (int)(d.rows() * d.cols());
  if (out && cap >= cnt) memcpy(out, d.data(), sizeof(float) * cnt);
  return cnt;
}
NAVH_API int navh_core_set_layer(void* h, const char* layer, const float* in) {
  Core* c = (Core*)h;
  if (!c->map.exists(layer)) c->map.add(layer);
  grid_map::Matrix& d = c->map[layer];
  memcpy(d.data(), in, sizeof(float) * d.rows() * d.cols());
  return 0;
}
// The cells a grid_map::LineIterator visits (LineIterator.cpp), for the known-answer tests: out = n * 2 ints.
NAVH_API int navh_core_line(void* h, double sx, double sy, double ex, double ey, int* out, int cap) {
  Core* c = (Core*)h;
  int k = 0;
  for (grid_map::LineIterator it(c->map, grid_map::Position(sx, sy), grid_map::Position(ex, ey)); !it.isPastEnd();
       ++it) {
    if (k < cap) {
      out[2 * k] = (*it)(0);
      out[2 * k + 1] = (*it)(1);
    }
    ++k;
  }
  return k;
}
// MapGlobalPlanner::ifBlocked's test (map_global_planner.h:39-54) with the reference's CircleIterator.
NAVH_API int navh_core_blocked(void* h, const char* layer, double x, double y, double radius) {
  Core* c = (Core*)h;
  grid_map::Position center(x, y);
  for (grid_map::CircleIterator it(c->map, center, radius); !it.isPastEnd(); ++it) {
    const float v = c->map.at(layer, *it);
    if (!std::isnan(v) && v > 0) return 1;
  }
  return 0;
}
#endif

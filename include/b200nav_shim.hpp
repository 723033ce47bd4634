/*
 * b200nav_shim.hpp -- header-only C++ host side with the reference's class interfaces, forwarding to the C ABI
 * (include/b200nav.h).  These are the classes a maintainer of jmloveyj/ros_navigation swaps in:
 *
 *   b200nav::VFH              same public surface as move_control::VFH
 *                             (move_control/include/move_control/vfh.h:182-245): the 19-argument constructor,
 *                             Init(), Update_VFH(double[361][2], int, float, float, float, int&, int&), the getters /
 *                             setters and the public float* Hist / OriginHist read by Steerer::pubHist
 *                             (move_control/src/steerer.cpp:201-220).  Added (non-breaking):
 *                             Update_VFH_FromGrid(...) = Steerer::getRangesFromSubmap + Update_VFH on the device.
 *   b200nav::GridLayers       the device-resident layers of one grid_map::GridMap (MapProvider's map_).
 *   b200nav::MapUpdater /     same surface as move_control::MapUpdater / LaserMapUpdater
 *   b200nav::LaserMapUpdater  (map_updater.h:8-80, laser_map_updater.h:13-43) minus ROS: the message intake keeps
 *                             running on the host and pushes RangeSamples; updateMap(minX,minY,maxX,maxY) drains the
 *                             buffer in order into its layer on the device.
 *
 * Error behaviour mirrors the reference: no exceptions from the update calls; failures are reported through
 * last_error() and leave outputs untouched (the reference only ROS_WARNs).  Construction failures throw
 * std::runtime_error (the reference would crash in the constructor as well).
 */
#ifndef B200NAV_SHIM_HPP
#define B200NAV_SHIM_HPP

#include <sys/time.h>

#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "b200nav.h"

namespace b200nav {

/* One CUDA device + stream; shared by the grid and the VFH object of a node. */
class Context {
 public:
  explicit Context(int device = 0, void* cuda_stream = nullptr) {
    if (b200nav_ctx_create(device, cuda_stream, &ctx_) != B200NAV_OK)
      throw std::runtime_error(std::string("b200nav_ctx_create: ") + b200nav_last_error(nullptr));
  }
  ~Context() { b200nav_ctx_destroy(ctx_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  b200nav_ctx* get() const { return ctx_; }
  const char* last_error() const { return b200nav_last_error(ctx_); }

 private:
  b200nav_ctx* ctx_ = nullptr;
};

/* Device-resident layers of one grid_map::GridMap (geometry as GridMap::setGeometry, GridMap.cpp:51-70). */
class GridLayers {
 public:
  GridLayers(Context& ctx, double length_x, double length_y, double resolution, double pos_x = 0.0, double pos_y = 0.0,
             int n_robots = 1)
      : ctx_(ctx) {
    if (b200nav_grid_create(ctx.get(), length_x, length_y, resolution, pos_x, pos_y, n_robots, &grid_) != B200NAV_OK)
      throw std::runtime_error(std::string("b200nav_grid_create: ") + ctx.last_error());
    b200nav_grid_size(grid_, &rows_, &cols_, &n_robots_);
  }
  ~GridLayers() { b200nav_grid_destroy(grid_); }
  GridLayers(const GridLayers&) = delete;
  GridLayers& operator=(const GridLayers&) = delete;

  bool exists(const std::string& layer) { return b200nav_grid_has_layer(grid_, layer.c_str()) != 0; }
  void add(const std::string& layer) { b200nav_grid_add_layer(grid_, layer.c_str()); }
  /* map_["master"] = map_["laser"] (map_provider.cpp:221): copying and zero-copy forms */
  int copyLayer(const std::string& dst, const std::string& src) {
    return b200nav_grid_copy_layer(grid_, dst.c_str(), src.c_str());
  }
  int aliasLayer(const std::string& alias, const std::string& target) {
    return b200nav_grid_alias_layer(grid_, alias.c_str(), target.c_str());
  }
  /* the two-layer compose the reference carries commented out (map_provider.cpp:218-220) */
  int composeMaster(const std::string& dst = "master", const std::string& range_layer = "range",
                    const std::string& laser_layer = "laser") {
    return b200nav_grid_compose_master(grid_, dst.c_str(), range_layer.c_str(), laser_layer.c_str());
  }
  /* true while the layer lives on the device as one byte per cell (see b200nav_grid_layer_format) */
  bool isCoded(const std::string& layer) { return b200nav_grid_layer_format(grid_, layer.c_str()) == B200NAV_LAYER_CODED; }
  /* Eigen::MatrixXf-compatible transfer: column-major rows x cols floats (grid_map::Matrix::data()) */
  int upload(const std::string& layer, const float* colmajor, int robot = 0) {
    return b200nav_grid_upload(grid_, robot, layer.c_str(), colmajor);
  }
  int download(const std::string& layer, float* colmajor, int robot = 0) {
    return b200nav_grid_download(grid_, robot, layer.c_str(), colmajor);
  }
  bool move(double x, double y, int robot = 0) { /* GridMap::move */
    int moved = 0;
    b200nav_grid_move(grid_, robot, x, y, &moved);
    return moved != 0;
  }
  int toOccupancyGrid(const std::string& layer, float data_min, float data_max, int8_t* out, int robot = 0) {
    return b200nav_grid_to_occupancy(grid_, robot, layer.c_str(), data_min, data_max, out);
  }
  /* MapGlobalPlanner::ifBlocked (map_global_planner.h:39-54) against the device layer */
  bool ifBlocked(double x, double y, double radius = 0.3, const std::string& layer = "master", int robot = 0) {
    const double xy[2] = {x, y};
    uint8_t out = 0;
    b200nav_grid_query_blocked(grid_, robot, layer.c_str(), xy, 1, radius, &out);
    return out != 0;
  }
  int rows() const { return rows_; }
  int cols() const { return cols_; }
  b200nav_grid* get() const { return grid_; }
  Context& context() const { return ctx_; }

 private:
  Context& ctx_;
  b200nav_grid* grid_ = nullptr;
  int rows_ = 0, cols_ = 0, n_robots_ = 1;
};

/* move_control::MapUpdater (map_updater.h:8-80). */
class MapUpdater {
 public:
  /* RangeSample (map_updater.h:28-32) == b200nav_sample */
  typedef b200nav_sample RangeSample;

  MapUpdater(GridLayers& map, const std::string& typeName, int robot = 0)
      : typeName_(typeName), map_(map), robot_(robot) {
    if (!map_.exists(typeName)) map_.add(typeName); /* map_updater.h:12-13 */
  }
  virtual ~MapUpdater() {}
  /* update map and point out the map range updated (map_updater.h:17) */
  virtual void updateMap(double& minX, double& minY, double& maxX, double& maxY) = 0;
  std::string getTypeName() { return typeName_; }

 protected:
  std::string typeName_;
  GridLayers& map_;
  int robot_;
};

/* move_control::LaserMapUpdater (laser_map_updater.h:13-43, laser_map_updater.cpp:7-21). */
class LaserMapUpdater : public MapUpdater {
 public:
  LaserMapUpdater(GridLayers& map, const std::string& sensorType = "laser", int robot = 0)
      : MapUpdater(map, sensorType, robot) {}

  /* What bufferIncomingMsg produces per cloud point (laser_map_updater.cpp:53-70); thread-safe like the reference
   * (laserSampleBufferMutex_). */
  void pushSample(double sx, double sy, double ex, double ey, bool ifClearEnd) {
    RangeSample s;
    s.sx = sx;
    s.sy = sy;
    s.ex = ex;
    s.ey = ey;
    s.clear_end = ifClearEnd ? 1 : 0;
    s.reserved = 0;
    std::lock_guard<std::mutex> lock(mutex_);
    buffer_.push_back(s);
  }
  void pushSamples(const RangeSample* s, size_t n) {
    std::lock_guard<std::mutex> lock(mutex_);
    buffer_.insert(buffer_.end(), s, s + n);
  }

  void updateMap(double& minX, double& minY, double& maxX, double& maxY) override {
    std::vector<RangeSample> copy;
    {
      std::lock_guard<std::mutex> lock(mutex_); /* laser_map_updater.cpp:9-13 */
      copy.swap(buffer_);
    }
    if (copy.empty()) return;
    double bbox[4] = {minX, minY, maxX, maxY};
    if (b200nav_himm_update(map_.get(), robot_, typeName_.c_str(), copy.data(), (int)copy.size(), bbox) == B200NAV_OK) {
      minX = bbox[0];
      minY = bbox[1];
      maxX = bbox[2];
      maxY = bbox[3];
    }
  }

 private:
  std::vector<RangeSample> buffer_;
  std::mutex mutex_;
};

/* move_control::VFH (vfh.h:182-245). */
class VFH {
 public:
  VFH(Context& ctx, double cell_size, int window_diameter, int sector_angle, double safety_dist_0ms,
      double safety_dist_1ms, int max_speed, int max_speed_narrow_opening, int max_speed_wide_opening,
      int max_acceleration, int min_turnrate, int max_turnrate_0ms, int max_turnrate_1ms,
      double min_turn_radius_safety_factor, double free_space_cutoff_0ms, double obs_cutoff_0ms,
      double free_space_cutoff_1ms, double obs_cutoff_1ms, double weight_desired_dir, double weight_current_dir)
      : Hist(nullptr), OriginHist(nullptr), ctx_(ctx) {
    b200nav_vfh_default_params(&p_);
    p_.cell_size = cell_size;
    p_.window_diameter = window_diameter;
    p_.sector_angle = sector_angle;
    p_.safety_dist_0ms = safety_dist_0ms;
    p_.safety_dist_1ms = safety_dist_1ms;
    p_.max_speed = max_speed;
    p_.max_speed_narrow_opening = max_speed_narrow_opening;
    p_.max_speed_wide_opening = max_speed_wide_opening;
    p_.max_acceleration = max_acceleration;
    p_.min_turnrate = min_turnrate;
    p_.max_turnrate_0ms = max_turnrate_0ms;
    p_.max_turnrate_1ms = max_turnrate_1ms;
    p_.min_turn_radius_safety_factor = min_turn_radius_safety_factor;
    p_.free_space_cutoff_0ms = free_space_cutoff_0ms;
    p_.obs_cutoff_0ms = obs_cutoff_0ms;
    p_.free_space_cutoff_1ms = free_space_cutoff_1ms;
    p_.obs_cutoff_1ms = obs_cutoff_1ms;
    p_.weight_desired_dir = weight_desired_dir;
    p_.weight_current_dir = weight_current_dir;
    gettimeofday(&last_update_time_, 0);
  }
  ~VFH() { b200nav_vfh_destroy(vfh_); }
  VFH(const VFH&) = delete;
  VFH& operator=(const VFH&) = delete;

  /* vfh.cpp:237-416: builds the tables (host) and uploads them.  Returns 1 like the reference, 0 on failure. */
  int Init() {
    if (vfh_) b200nav_vfh_destroy(vfh_);
    vfh_ = nullptr;
    if (b200nav_vfh_create(ctx_.get(), &p_, 1, &vfh_) != B200NAV_OK) return 0;
    hist_size_ = b200nav_vfh_hist_size(vfh_);
    hist_.assign(hist_size_, 0.f);
    origin_.assign(hist_size_, 0.f);
    Hist = hist_.data();
    OriginHist = origin_.data();
    gettimeofday(&last_update_time_, 0); /* vfh.cpp:413 */
    return 1;
  }

  /* vfh.cpp:480-605.  Returns 1 like the reference. */
  int Update_VFH(double laser_ranges[361][2], int current_speed, float goal_direction, float goal_distance,
                 float goal_distance_tolerance, int& chosen_speed, int& chosen_turnrate) {
    b200nav_vfh_input in = make_input(current_speed, goal_direction, goal_distance, goal_distance_tolerance);
    b200nav_command out;
    if (vfh_ && b200nav_vfh_update_ranges(vfh_, 0, &laser_ranges[0][0], &in, &out) == B200NAV_OK)
      finish(out, chosen_speed, chosen_turnrate);
    return 1;
  }

  /* Steerer::getRangesFromSubmap (steerer.cpp:147-191) + Update_VFH, reading `layer` of the device grid at the robot
   * pose (x, y, yaw); saves the host submap copy and the 961 atan2 calls. */
  int Update_VFH_FromGrid(GridLayers& map, const std::string& layer, double x, double y, double yaw, int current_speed,
                          float goal_direction, float goal_distance, float goal_distance_tolerance, int& chosen_speed,
                          int& chosen_turnrate) {
    return update_from_device_grid(map.get(), layer, x, y, yaw, current_speed, goal_direction, goal_distance,
                                   goal_distance_tolerance, chosen_speed, chosen_turnrate);
  }
  int update_from_device_grid(b200nav_grid* grid, const std::string& layer, double x, double y, double yaw,
                              int current_speed, float goal_direction, float goal_distance,
                              float goal_distance_tolerance, int& chosen_speed, int& chosen_turnrate) {
    b200nav_vfh_input in = make_input(current_speed, goal_direction, goal_distance, goal_distance_tolerance);
    in.x = x;
    in.y = y;
    in.yaw = yaw;
    b200nav_command out;
    if (vfh_ && b200nav_vfh_update_grid(vfh_, grid, layer.c_str(), 0, &in, &out) == B200NAV_OK)
      finish(out, chosen_speed, chosen_turnrate);
    return 1;
  }

  /* Get methods (vfh.h:224-233) */
  int GetMinTurnrate() { return p_.min_turnrate; }
  float GetDesiredAngle() { return desired_angle_; }
  float GetPickedAngle() { return picked_angle_; }
  int GetMaxTurnrate(int speed) { return vfh_ ? b200nav_vfh_get_max_turnrate(vfh_, speed) : 0; }
  int GetCurrentMaxSpeed() { return current_max_speed_ < 0 ? p_.max_speed : current_max_speed_; }
  /* Set methods (vfh.h:236-238) */
  void SetRobotRadius(float robot_radius) { p_.robot_radius = robot_radius; } /* before Init(), like the reference */
  void SetMinTurnrate(int min_turnrate) { p_.min_turnrate = min_turnrate; }
  void SetCurrentMaxSpeed(int max_speed) {
    current_max_speed_ = max_speed < p_.max_speed ? max_speed : p_.max_speed;
    if (vfh_) b200nav_vfh_set_current_max_speed(vfh_, max_speed);
  }
  int getHistSize() { return hist_size_; }
  int getSectorAngle() { return p_.sector_angle; }

  /* Deterministic clock for tests: seconds to report as elapsed on the next update instead of gettimeofday. */
  void SetNextElapsed(double dt) {
    forced_dt_ = dt;
    has_forced_dt_ = true;
  }
  b200nav_vfh* get() const { return vfh_; }

  /* The Histogram.  Public so that monitoring tools can get at it (vfh.h:240-244). */
  float* Hist;
  float* OriginHist;

 private:
  b200nav_vfh_input make_input(int current_speed, float gd, float gdist, float tol) {
    b200nav_vfh_input in;
    in.x = in.y = in.yaw = 0.0;
    in.current_speed = current_speed;
    in.goal_direction = gd;
    in.goal_distance = gdist;
    in.goal_tolerance = tol;
    desired_angle_ = gd;
    /* vfh.cpp:521-531: elapsed wall time between updates (TIMESUB) */
    timeval now;
    gettimeofday(&now, 0);
    long sec = now.tv_sec - last_update_time_.tv_sec, usec = now.tv_usec - last_update_time_.tv_usec;
    if (usec < 0) {
      sec -= 1;
      usec += 1000000;
    }
    in.dt = sec + ((double)usec / 1000000);
    last_update_time_ = now;
    if (has_forced_dt_) {
      in.dt = forced_dt_;
      has_forced_dt_ = false;
    }
    return in;
  }
  void finish(const b200nav_command& out, int& chosen_speed, int& chosen_turnrate) {
    chosen_speed = out.speed;
    chosen_turnrate = out.turnrate;
    picked_angle_ = out.picked_angle;
    b200nav_vfh_read_state(vfh_, 0, OriginHist, Hist, nullptr, nullptr, nullptr);
  }

  Context& ctx_;
  b200nav_vfh_params p_;
  b200nav_vfh* vfh_ = nullptr;
  int hist_size_ = 0;
  std::vector<float> hist_, origin_;
  float desired_angle_ = 90.f, picked_angle_ = 90.f;
  int current_max_speed_ = -1;
  timeval last_update_time_;
  double forced_dt_ = 0.0;
  bool has_forced_dt_ = false;
};

}  // namespace b200nav
#endif

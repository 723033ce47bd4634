/*
 * b200nav_dropin.hpp -- what the drop-in class headers of include/move_control/ share: the process-wide device
 * context and the device twin of a host grid_map::GridMap.
 *
 * The reference's nodes construct their mapping / steering classes without any device argument
 *   new LaserMapUpdater(nh_, tf_, map_, sensorType)        move_control/src/map_provider.cpp:12-15,262-266
 *   new VFH(cell_size, ..., weight_current_dir)            move_control/src/steerer.cpp:122-129
 * so the drop-in classes take the CUDA context from here instead of from a constructor parameter:
 * b200nav::default_context() creates one b200nav_ctx per process on device $B200NAV_DEVICE (default 0).
 *
 * grid_map::GridMap stays the host-side owner of the map (MapProvider composes, moves, publishes and cuts submaps
 * out of it: map_provider.cpp:93-100,177-188,207-223).  b200nav::DeviceMap is its twin in HBM:
 *   - created on first use with the GridMap's geometry (GridMap::setGeometry, grid_map_core/src/GridMap.cpp:51-70);
 *   - before an update, if the host map was moved or re-centred since the last one (GridMap::move,
 *     GridMap.cpp:346-412, run by MapProvider::loopMoveMap), the layer is re-uploaded together with the new
 *     position / start index - the host copy was in step before the move, so it is the truth after it;
 *   - after an update the layer is copied back into the host matrix, so every host-side reader
 *     (composeMasterMapFromLayerdMap, getSubmap, toOccupancyGrid, the planners) keeps working unchanged.
 * Consumers that want the data to stay in HBM (VFH::Update_VFH_FromGrid, the batched C ABI) skip the copy-back.
 */
#ifndef B200NAV_DROPIN_HPP
#define B200NAV_DROPIN_HPP

#include <stdlib.h>

#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "b200nav.h"
#include "b200nav_shim.hpp"
#include "grid_map_core/grid_map_core.hpp"

namespace b200nav {

/* One context per process, created on first use.  Throws std::runtime_error without a CUDA device: there is no CPU
 * fallback behind these classes. */
inline Context& default_context() {
  static std::unique_ptr<Context> ctx;
  static std::mutex m;
  std::lock_guard<std::mutex> lock(m);
  if (!ctx) {
    const char* dev = getenv("B200NAV_DEVICE");
    ctx.reset(new Context(dev ? atoi(dev) : 0));
  }
  return *ctx;
}

class DeviceMap {
 public:
  explicit DeviceMap(grid_map::GridMap& host) : host_(host) {}
  ~DeviceMap() {
    if (grid_) b200nav_grid_destroy(grid_);
  }
  DeviceMap(const DeviceMap&) = delete;
  DeviceMap& operator=(const DeviceMap&) = delete;

  b200nav_grid* grid() { return ensure() ? grid_ : nullptr; }

  /* Apply `n` samples in order to `layer` on the device and bring the host matrix up to date. */
  bool update(const std::string& layer, const b200nav_sample* samples, int n, double* bbox, bool copy_back = true) {
    if (!make_current(layer, true)) return false;
    if (n > 0 && b200nav_himm_update(grid_, 0, layer.c_str(), samples, n, bbox) != B200NAV_OK) return false;
    if (n > 0 && copy_back) return pull(layer);
    return true;
  }
  /* Make the device copy of `layer` usable by a device-side reader (VFH::Update_VFH_FromGrid).  Layers that an
   * updater maintains through update() are already current; any other layer (e.g. "master", which MapProvider
   * composes on the host) is uploaded on every call. */
  bool make_readable(const std::string& layer) { return make_current(layer, false); }
  /* Forget what is known about the device copy of `layer` (the next update uploads the host matrix first).  Called
   * when an updater finds its layer missing on the host: a GridMap that was just (re)created - possibly at the address
   * of an earlier one - must not inherit that one's device state. */
  void invalidate(const std::string& layer) { epoch_of_.erase(layer); }
  /* device -> host matrix */
  bool pull(const std::string& layer) {
    grid_map::Matrix& m = host_[layer];
    return b200nav_grid_download(grid_, 0, layer.c_str(), m.data()) == B200NAV_OK;
  }
  const char* last_error() { return default_context().last_error(); }

 private:
  struct Stamp {
    double px, py;
    int s0, s1;
    bool operator==(const Stamp& o) const { return px == o.px && py == o.py && s0 == o.s0 && s1 == o.s1; }
  };
  Stamp stamp() const {
    return Stamp{host_.getPosition()(0), host_.getPosition()(1), host_.getStartIndex()(0), host_.getStartIndex()(1)};
  }
  /* (re)create the device grid when the host geometry (size, resolution) changed */
  bool ensure() {
    const int rows = host_.getSize()(0), cols = host_.getSize()(1);
    const double res = host_.getResolution();
    if (grid_ && rows == rows_ && cols == cols_ && res == res_) return true;
    if (grid_) b200nav_grid_destroy(grid_);
    grid_ = nullptr;
    epoch_of_.clear();
    observed_.px = std::numeric_limits<double>::quiet_NaN(); /* first use sets the geometry */
    if (rows <= 0 || cols <= 0) return false;
    if (b200nav_grid_create(default_context().get(), host_.getLength()(0), host_.getLength()(1), res,
                            host_.getPosition()(0), host_.getPosition()(1), 1, &grid_) != B200NAV_OK)
      return false;
    int r = 0, c = 0, n = 0;
    b200nav_grid_size(grid_, &r, &c, &n);
    if (r != rows || c != cols) { /* cannot happen: same rounding as GridMap::setGeometry */
      b200nav_grid_destroy(grid_);
      grid_ = nullptr;
      return false;
    }
    rows_ = rows;
    cols_ = cols;
    res_ = res;
    return true;
  }
  /* host matrix -> device */
  bool push(const std::string& layer) {
    if (!b200nav_grid_has_layer(grid_, layer.c_str())) b200nav_grid_add_layer(grid_, layer.c_str());
    const grid_map::Matrix& m = host_[layer];
    return b200nav_grid_upload(grid_, 0, layer.c_str(), m.data()) == B200NAV_OK;
  }
  /* The host map is only ever moved as a whole (GridMap::move clears the dropped strips of every layer on the host),
   * so "the host was moved since I last looked" is visible in its position / start index.  Every observed change
   * starts a new epoch; a device-maintained layer is current iff it was pushed or updated in this epoch.  (A map
   * that moves away and exactly back between two updates is not noticed - at the reference's 2 Hz move and 5 Hz
   * update rates that needs two moves within one update period.) */
  bool make_current(const std::string& layer, bool maintained_here) {
    if (!ensure()) return false;
    const Stamp now = stamp();
    if (!(now == observed_)) {
      if (b200nav_grid_set_geometry(grid_, 0, now.px, now.py, now.s0, now.s1) != B200NAV_OK) return false;
      observed_ = now;
      epoch_ += 1;
    }
    auto it = epoch_of_.find(layer);
    const bool current = it != epoch_of_.end() && it->second == epoch_;
    if (maintained_here) {
      if (!current && !push(layer)) return false;
      epoch_of_[layer] = epoch_;
      return true;
    }
    if (current) return true; /* an updater keeps this layer current on the device */
    return push(layer);       /* host-composed layer: no way to know when the host rewrote it */
  }

  grid_map::GridMap& host_;
  b200nav_grid* grid_ = nullptr;
  int rows_ = 0, cols_ = 0;
  double res_ = 0.0;
  Stamp observed_ = {std::numeric_limits<double>::quiet_NaN(), 0.0, 0, 0};
  long epoch_ = 0;
  std::map<std::string, long> epoch_of_;
};

/* The device twin of `host` (one per GridMap object; MapProvider's map_ lives as long as its updaters). */
inline DeviceMap& device_map_for(grid_map::GridMap& host) {
  default_context(); /* constructed before (and so destroyed after) the twins that hold grids of it */
  static std::map<grid_map::GridMap*, std::unique_ptr<DeviceMap>> twins;
  static std::mutex m;
  std::lock_guard<std::mutex> lock(m);
  std::unique_ptr<DeviceMap>& t = twins[&host];
  if (!t) t.reset(new DeviceMap(host));
  return *t;
}

}  // namespace b200nav
#endif

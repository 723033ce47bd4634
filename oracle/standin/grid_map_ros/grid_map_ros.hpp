// Stand-in for grid_map_ros (its real converter needs cv_bridge, rosbag and the grid_map message packages).  Only the
// one function MapProvider calls exists.  This is a STAND-IN RESTATEMENT of GridMapRosConverter::toOccupancyGrid
// (grid_map_ros/src/GridMapRosConverter.cpp:251-287), written over the reference's real GridMap API; it does not pin
// anything - the occupancy conversion is pinned by oracle/himm_oracle.cpp's own restatement only.
#ifndef B200NAV_STANDIN_GRID_MAP_ROS_HPP
#define B200NAV_STANDIN_GRID_MAP_ROS_HPP
#include <grid_map_core/grid_map_core.hpp>

#include "standin_ros_core.hpp"
namespace grid_map {
struct GridMapRosConverter {
  static void toOccupancyGrid(const grid_map::GridMap& gridMap, const std::string& layer, float dataMin,
                              float dataMax, nav_msgs::OccupancyGrid& out) {
    out.header.frame_id = gridMap.getFrameId();
    out.header.stamp.fromNSec(gridMap.getTimestamp());
    out.info.map_load_time = out.header.stamp;
    out.info.resolution = gridMap.getResolution();
    out.info.width = gridMap.getSize()(0);
    out.info.height = gridMap.getSize()(1);
    out.info.origin.position.x = gridMap.getPosition().x() - 0.5 * gridMap.getLength().x();
    out.info.origin.position.y = gridMap.getPosition().y() - 0.5 * gridMap.getLength().y();
    out.info.origin.orientation.w = 1.0;
    const size_t n = (size_t)gridMap.getSize()(0) * gridMap.getSize()(1);
    out.data.resize(n);
    for (GridMapIterator it(gridMap); !it.isPastEnd(); ++it) {
      float v = (gridMap.at(layer, *it) - dataMin) / (dataMax - dataMin);
      if (std::isnan(v) || v < 0)
        v = -1;
      else
        v = 0.f + std::min(std::max(0.0f, v), 1.0f) * 100.f;
      const size_t lin = getLinearIndexFromIndex(it.getUnwrappedIndex(), gridMap.getSize(), false);
      out.data[n - lin - 1] = (int8_t)v;
    }
  }
};
}  // namespace grid_map
#endif

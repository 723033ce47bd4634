/*
 * move_control/map_updater.h -- drop-in replacement for the reference header of the same name
 * (move_control/include/move_control/map_updater.h:8-80).  Same namespace, class name, constructor, virtual
 * interface and protected members, so that MapProvider's factory (move_control/src/map_provider.cpp:12-15,262-266),
 * its updater registry (:235-260) and user subclasses compile unchanged; the HIMM arithmetic itself
 * (lineOnMap / clearCell / markCell, :38-71, and grid_map::LineIterator behind it) runs on the GPU through the C ABI
 * (b200nav_himm_update, include/b200nav.h).  The host grid_map::GridMap remains the owner of the map; its device
 * twin is described in b200nav_dropin.hpp.
 */
#ifndef MAP_UPDATER_H
#define MAP_UPDATER_H
#include "ros/ros.h"
#include <tf/transform_listener.h>
#include "grid_map_core/grid_map_core.hpp"

#include <algorithm>
#include <string>
#include <vector>

#include "../b200nav_dropin.hpp"

namespace move_control {

class MapUpdater {
public:
    MapUpdater(ros::NodeHandle& nh, tf::TransformListener& tf, grid_map::GridMap &map, const std::string& typeName):
        typeName_(typeName), nh_(nh), tf_(tf), map_(map), device_(b200nav::device_map_for(map)) {
        if (!map_.exists(typeName)) {
            map_.add(typeName);
            device_.invalidate(typeName);   // a fresh host layer: whatever the device twin held under this name is stale
        }
    }
    virtual ~MapUpdater() {}
    // update map and point out the map range updated
    virtual void updateMap(double &minX, double &minY, double &maxX, double &maxY) = 0;

    std::string getTypeName() {
        return typeName_;
    }

    virtual void addMonitorTopic(const std::string &topicName) = 0;

protected:
    std::string typeName_;
    typedef struct {
      grid_map::Position start;
      grid_map::Position end;
      bool ifClearEnd;   // true: clear the end in the map. false: mark as obstacle
    } RangeSample;

    ros::NodeHandle& nh_;
    tf::TransformListener& tf_;
    grid_map::GridMap& map_;

    // One ray (reference: map_updater.h:38-50).  Subclasses that drain a whole buffer should call applySamples.
    void lineOnMap(const RangeSample &rangeSample) {
        applySamples(&rangeSample, 1, nullptr);
    }

    // The body of Laser/RangeMapUpdater::updateMap (laser_map_updater.cpp:15-20): every sample in order, one device
    // call, bounding box like touch().  bbox = {minX, minY, maxX, maxY} or NULL.
    bool applySamples(const RangeSample* samples, size_t n, double* bbox) {
        staging_.resize(n);
        for (size_t i = 0; i < n; ++i) {
            b200nav_sample& s = staging_[i];
            s.sx = samples[i].start(0);
            s.sy = samples[i].start(1);
            s.ex = samples[i].end(0);
            s.ey = samples[i].end(1);
            s.clear_end = samples[i].ifClearEnd ? 1 : 0;
            s.reserved = 0;
        }
        return device_.update(typeName_, staging_.data(), (int)n, bbox);
    }

    void touch(double &minX, double &minY, double &maxX, double &maxY, double &x, double &y) {
         minX = std::min(minX, x);
         minY = std::min(minY, y);
         maxX = std::max(maxX, x);
         maxY = std::max(maxY, y);
    }

    b200nav::DeviceMap& device_;

private:
    std::vector<b200nav_sample> staging_;
};

}
#endif // MAP_UPDATER_H

/*
 * move_control/range_map_updater.h -- drop-in replacement for the reference's RangeMapUpdater (sonar rays into layer
 * "range": move_control/include/move_control/range_map_updater.h:11-36, move_control/src/range_map_updater.cpp).
 * Header-only; same intake as the reference (one ray per sensor_msgs/Range message: sensor origin -> the point
 * `range` metres along the sensor's x axis, both through tf), buffer drained on the GPU by updateMap().
 */
#include "move_control/map_updater.h"
#include "tf/message_filter.h"
#include "message_filters/subscriber.h"
#include "sensor_msgs/Range.h"
#include <boost/thread/mutex.hpp>
#include <boost/bind.hpp>
#include <boost/shared_ptr.hpp>

#ifndef RANGE_MAP_UPDATER_H
#define RANGE_MAP_UPDATER_H

namespace move_control {
class RangeMapUpdater: public MapUpdater {
public:
    RangeMapUpdater(ros::NodeHandle& nh, tf::TransformListener& tf, grid_map::GridMap& map, const std::string& sensorType="range"):
        MapUpdater(nh, tf, map, sensorType) {}

    ~RangeMapUpdater() {}

    // update map and point out the map range updated
    void updateMap(double &minX, double &minY, double &maxX, double &maxY) {
        std::vector<RangeSample> drained;
        {
            boost::unique_lock<boost::mutex> lock(bufferMutex_);
            drained.swap(buffer_);
        }
        if (drained.empty())
            return;
        double bbox[4] = {minX, minY, maxX, maxY};
        if (applySamples(drained.data(), drained.size(), bbox)) {
            minX = bbox[0];
            minY = bbox[1];
            maxX = bbox[2];
            maxY = bbox[3];
        } else {
            ROS_WARN("b200nav: HIMM update failed: %s", device_.last_error());
        }
    }

    std::string getTypeName() {
        return "range";
    }

    void addMonitorTopic(const std::string &topicName) {
        typedef message_filters::Subscriber<sensor_msgs::Range> Sub;
        typedef tf::MessageFilter<sensor_msgs::Range> Filter;
        boost::shared_ptr<Sub> sub(new Sub(nh_, topicName, 50));
        boost::shared_ptr<Filter> filter(new Filter(*sub, tf_, "/odom", 50));
        filter->registerCallback(boost::bind(&RangeMapUpdater::onRange, this, _1));
        subs_.push_back(sub);
        filters_.push_back(filter);
    }

private:
    std::vector<boost::shared_ptr<tf::MessageFilter<sensor_msgs::Range> > > filters_;
    std::vector<boost::shared_ptr<message_filters::Subscriber<sensor_msgs::Range> > > subs_;
    std::vector<RangeSample> buffer_;
    boost::mutex bufferMutex_;

    void onRange(const sensor_msgs::RangeConstPtr& msg) {
        const std::string mapFrame = map_.getFrameId();
        geometry_msgs::PointStamped onSensor, onMap;
        onSensor.header.stamp = msg->header.stamp;
        onSensor.header.frame_id = msg->header.frame_id;

        RangeSample sample;
        sample.ifClearEnd = !(msg->range < msg->max_range);   // a reading at max range only clears
        // Like the reference (range_map_updater.cpp:45-68) a failed lookup is only warned about: the sample is still
        // buffered with whatever the previous lookup left in `onMap`.
        try {
            tf_.transformPoint(mapFrame, onSensor, onMap);
        } catch (tf::TransformException &ex) {
            ROS_WARN("b200nav: %s", ex.what());
        }
        sample.start = grid_map::Position(onMap.point.x, onMap.point.y);
        onSensor.point.x = msg->range;
        try {
            tf_.transformPoint(mapFrame, onSensor, onMap);
        } catch (tf::TransformException &ex) {
            ROS_WARN("b200nav: %s", ex.what());
        }
        sample.end = grid_map::Position(onMap.point.x, onMap.point.y);

        boost::unique_lock<boost::mutex> lock(bufferMutex_);
        buffer_.push_back(sample);
    }
};
}

#endif // RANGE_MAP_UPDATER_H

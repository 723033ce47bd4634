/*
 * move_control/laser_map_updater.h -- drop-in replacement for the reference's LaserMapUpdater
 * (move_control/include/move_control/laser_map_updater.h:13-43, move_control/src/laser_map_updater.cpp).
 * Header-only: remove laser_map_updater.cpp from the catkin target and link libb200nav.so.
 *
 * The message intake stays on the host and on ROS, as in the reference: scans arrive through a tf::MessageFilter,
 * at most one per 0.2 s is accepted, scans finer than 0.017 rad are thinned to about one reading per degree,
 * laser_geometry projects into the map frame, and every cloud point becomes one RangeSample.  updateMap() then
 * hands the drained buffer to the GPU in ONE call (b200nav_himm_update) instead of walking it cell by cell.
 */
#include "move_control/map_updater.h"
#include "tf/message_filter.h"
#include "message_filters/subscriber.h"
#include "sensor_msgs/LaserScan.h"
#include "sensor_msgs/PointCloud2.h"
#include "sensor_msgs/point_cloud2_iterator.h"
#include "laser_geometry/laser_geometry.h"
#include <boost/thread/mutex.hpp>
#include <boost/bind.hpp>
#include <boost/shared_ptr.hpp>

#ifndef LASER_MAP_UPDATER_H
#define LASER_MAP_UPDATER_H
using namespace grid_map;   // the reference header exports this (laser_map_updater.h:10); map_provider.cpp relies on it

namespace move_control {
class LaserMapUpdater: public MapUpdater {
public:
    LaserMapUpdater(ros::NodeHandle& nh, tf::TransformListener& tf, grid_map::GridMap& map, const std::string& sensorType="laser"):
        MapUpdater(nh, tf, map, sensorType),
        lastAccepted_(0),
        acceptPeriod_(0.2) {}

    ~LaserMapUpdater() {}

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

    void addMonitorTopic(const std::string &topicName) {
        typedef message_filters::Subscriber<sensor_msgs::LaserScan> Sub;
        typedef tf::MessageFilter<sensor_msgs::LaserScan> Filter;
        boost::shared_ptr<Sub> sub(new Sub(nh_, topicName, 50));
        boost::shared_ptr<Filter> filter(new Filter(*sub, tf_, "/odom", 50));
        filter->registerCallback(boost::bind(&LaserMapUpdater::onScan, this, _1));
        subs_.push_back(sub);
        filters_.push_back(filter);
    }

private:
    std::vector<boost::shared_ptr<tf::MessageFilter<sensor_msgs::LaserScan> > > filters_;
    std::vector<boost::shared_ptr<message_filters::Subscriber<sensor_msgs::LaserScan> > > subs_;
    std::vector<RangeSample> buffer_;
    boost::mutex bufferMutex_;
    ros::Time lastAccepted_;
    ros::Duration acceptPeriod_;

    // one reading per >= 0.017 rad (reference: simplifyLaserScan): ranges[0], then every ranges[i] at which the float
    // accumulator of angle_increment reaches 0.017; the thinned scan's angle_increment is the last accumulated value
    static void thin(const sensor_msgs::LaserScan& in, sensor_msgs::LaserScan& out) {
        out.header = in.header;
        out.angle_min = in.angle_min;
        out.angle_max = in.angle_max;
        out.range_min = in.range_min;
        out.range_max = in.range_max;
        out.time_increment = in.time_increment;
        out.scan_time = in.scan_time;
        out.ranges.clear();
        if (in.ranges.empty())
            return;
        out.ranges.push_back(in.ranges[0]);
        float acc = 0.0;
        for (size_t i = 0; i < in.ranges.size(); ++i) {
            acc += in.angle_increment;
            if (acc >= 0.017) {
                out.angle_increment = acc;
                acc = 0.0;
                out.ranges.push_back(in.ranges[i]);
            }
        }
    }

    void onScan(const sensor_msgs::LaserScanConstPtr& msg) {
        const ros::Time now = ros::Time::now();
        if (lastAccepted_ + acceptPeriod_ > now)
            return;   // too soon after the previous accepted scan
        lastAccepted_ = now;

        const std::string mapFrame = map_.getFrameId();
        RangeSample sample;
        sensor_msgs::PointCloud2 cloud;
        try {
            // the sensor's origin in the map frame at the scan's stamp
            geometry_msgs::PointStamped sensorOrigin, originOnMap;
            sensorOrigin.header.stamp = msg->header.stamp;
            sensorOrigin.header.frame_id = msg->header.frame_id;
            tf_.transformPoint(mapFrame, sensorOrigin, originOnMap);
            sample.start = Position(originOnMap.point.x, originOnMap.point.y);

            laser_geometry::LaserProjection projector;
            if (msg->angle_increment < 0.017) {
                sensor_msgs::LaserScan thinned;
                thin(*msg, thinned);
                projector.transformLaserScanToPointCloud(mapFrame, thinned, cloud, tf_);
            } else {
                projector.transformLaserScanToPointCloud(mapFrame, *msg, cloud, tf_);
            }
        } catch (tf::TransformException &ex) {
            ROS_WARN("b200nav: scan dropped, no transform to %s: %s", mapFrame.c_str(), ex.what());
            return;
        }

        sensor_msgs::PointCloud2Iterator<float> x(cloud, "x"), y(cloud, "y");
        sensor_msgs::PointCloud2Iterator<int> index(cloud, "index");
        boost::unique_lock<boost::mutex> lock(bufferMutex_);
        for (; x != x.end(); ++x, ++y, ++index) {
            sample.end = Position(*x, *y);
            // Kept from the reference on purpose (laser_map_updater.cpp:62-66): `index` counts readings of the
            // PROJECTED scan but is looked up in the ORIGINAL one, so after thinning the flag follows a different beam.
            const float looked_up = msg->ranges[*index];
            sample.ifClearEnd = std::isinf(looked_up) || looked_up == msg->range_max;
            buffer_.push_back(sample);
        }
    }
};
}

#endif // LASER_MAP_UPDATER_H

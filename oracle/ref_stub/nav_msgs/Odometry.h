// nav_msgs/Odometry.h — STUB (oracle/_ref)
#pragma once
#include "sensor_msgs/Imu.h"
namespace nav_msgs {
struct Odometry {
    std_msgs::Header header;
    geometry_msgs::PoseWithCovariance pose;
    typedef std::shared_ptr<const Odometry> ConstPtr;
};
}  // namespace nav_msgs

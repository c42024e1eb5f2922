// sensor_msgs/Imu.h — STUB (oracle/_ref)
#pragma once
#include <memory>
#include "ros/ros.h"
namespace geometry_msgs {
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Point { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Pose { Point position; Quaternion orientation; };
struct PoseWithCovariance { Pose pose; };
}  // namespace geometry_msgs
namespace sensor_msgs {
struct Imu {
    std_msgs::Header header;
    geometry_msgs::Quaternion orientation;
    geometry_msgs::Vector3 angular_velocity, linear_acceleration;
    typedef std::shared_ptr<const Imu> ConstPtr;
};
struct Image { std_msgs::Header header; };
typedef std::shared_ptr<const Image> ImageConstPtr;
}  // namespace sensor_msgs

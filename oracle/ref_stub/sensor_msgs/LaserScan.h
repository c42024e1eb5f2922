// sensor_msgs/LaserScan.h — STUB (oracle/_ref): field names and types of the ROS message (float32 angles / ranges).
#pragma once
#include <memory>
#include <vector>
#include "ros/ros.h"
namespace sensor_msgs {
struct LaserScan {
    std_msgs::Header header;
    float angle_min = 0, angle_max = 0, angle_increment = 0, time_increment = 0, scan_time = 0, range_min = 0, range_max = 0;
    std::vector<float> ranges, intensities;
    typedef std::shared_ptr<LaserScan> Ptr;
    typedef std::shared_ptr<const LaserScan> ConstPtr;
};
typedef std::shared_ptr<const LaserScan> LaserScanConstPtr;
}  // namespace sensor_msgs

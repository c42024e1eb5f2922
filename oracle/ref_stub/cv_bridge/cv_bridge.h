// cv_bridge/cv_bridge.h — STUB (oracle/_ref): sensor::camera (src/trajectory/sensor.h:128-147) only has to compile.
#pragma once
#include <memory>
#include <string>
#include "opencv2/opencv.hpp"
#include "sensor_msgs/Imu.h"
namespace cv_bridge {
struct CvImage { cv::Mat image; };
typedef std::shared_ptr<const CvImage> CvImageConstPtr;
inline CvImageConstPtr toCvShare(const sensor_msgs::ImageConstPtr&, const std::string&) { return std::make_shared<const CvImage>(); }
}  // namespace cv_bridge

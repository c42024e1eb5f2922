// opencv2/opencv.hpp — STUB (oracle/_ref): frame_info carries a cv::Mat the laser path never reads.
#pragma once
namespace cv {
class Mat {};
template <class T> struct Point_ { T x, y; Point_() : x(), y() {} Point_(T a, T b) : x(a), y(b) {} };
typedef Point_<float> Point2f;
typedef Point_<int> Point;
}  // namespace cv

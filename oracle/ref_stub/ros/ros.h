// ros/ros.h — STUB (test infrastructure, oracle/_ref): just enough for the reference headers to compile without ROS.
#pragma once
#include <cstdio>
#include <memory>
#include <string>
namespace ros {
struct Time {
    double sec_ = 0.0;
    Time() {}
    explicit Time(double s) : sec_(s) {}
    double toSec() const { return sec_; }
    static Time now() { return Time(); }
};
class NodeHandle {
public:
    NodeHandle() {}
    explicit NodeHandle(const std::string&) {}
    template <class T> bool getParam(const std::string&, T&) const { return false; }
};
}  // namespace ros
namespace std_msgs {
struct Header { unsigned seq = 0; ros::Time stamp; std::string frame_id; };
}
#define ROS_STUB_LOG(...) do { std::fprintf(stderr, __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)
#define ROS_INFO(...) do { } while (0)
#define ROS_WARN(...) ROS_STUB_LOG(__VA_ARGS__)
#define ROS_ERROR(...) ROS_STUB_LOG(__VA_ARGS__)

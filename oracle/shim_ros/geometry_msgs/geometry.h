// oracle/shim_ros — TEST INFRASTRUCTURE (see ros/ros.h): the geometry_msgs records the two nodes fill in.
#pragma once
#include <array>
#include <std_msgs/Header.h>
namespace geometry_msgs {
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Point { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 0; };
struct Pose { Point position; Quaternion orientation; };
struct PoseWithCovariance { Pose pose; std::array<double, 36> covariance{}; };
struct Twist { Vector3 linear, angular; };
struct TwistWithCovariance { Twist twist; std::array<double, 36> covariance{}; };
struct PoseStamped { std_msgs::Header header; Pose pose; };
}  // namespace geometry_msgs

// oracle/shim_ros — TEST INFRASTRUCTURE (see ros/ros.h).
#pragma once
#include <array>
#include <std_msgs/Header.h>
namespace sensor_msgs {
struct NavSatStatus { int8_t status = 0; uint16_t service = 0; };
struct NavSatFix { std_msgs::Header header; NavSatStatus status; double latitude = 0, longitude = 0, altitude = 0; std::array<double, 9> position_covariance{}; };
typedef std::shared_ptr<NavSatFix const> NavSatFixConstPtr;
}

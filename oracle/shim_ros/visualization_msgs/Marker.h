// oracle/shim_ros — TEST INFRASTRUCTURE (see ros/ros.h).
#pragma once
#include <geometry_msgs/geometry.h>
namespace visualization_msgs {
struct ColorRGBA { float r = 0, g = 0, b = 0, a = 0; };
struct Marker {
    enum { ARROW = 0, CUBE = 1, SPHERE = 2, CYLINDER = 3, LINE_STRIP = 4, LINE_LIST = 5, CUBE_LIST = 6, SPHERE_LIST = 7, ADD = 0 };
    std_msgs::Header header; std::string ns; int32_t id = 0, type = 0, action = 0; geometry_msgs::Pose pose; geometry_msgs::Vector3 scale; ColorRGBA color;
    std::vector<geometry_msgs::Point> points;
};
}

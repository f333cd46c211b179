// oracle/shim_ros — TEST INFRASTRUCTURE (see ros/ros.h).
#pragma once
#include <visualization_msgs/Marker.h>
namespace visualization_msgs { struct MarkerArray { std::vector<Marker> markers; }; }

// oracle/shim_ros — TEST INFRASTRUCTURE (see ros/ros.h).
#pragma once
#include <geometry_msgs/geometry.h>

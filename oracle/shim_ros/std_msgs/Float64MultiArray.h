// oracle/shim_ros — TEST INFRASTRUCTURE (see ros/ros.h).
#pragma once
#include <ros/ros.h>
namespace std_msgs { struct Float64MultiArray { std::vector<double> data; typedef std::shared_ptr<Float64MultiArray const> ConstPtr; }; }

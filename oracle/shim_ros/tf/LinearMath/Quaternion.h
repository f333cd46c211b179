// oracle/shim_ros — TEST INFRASTRUCTURE (see ros/ros.h).
#pragma once
#include <tf/transform_datatypes.h>

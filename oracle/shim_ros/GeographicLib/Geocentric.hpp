// oracle/shim_ros — TEST INFRASTRUCTURE (see ros/ros.h).
#pragma once

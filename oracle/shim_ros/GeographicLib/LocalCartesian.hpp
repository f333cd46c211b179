// oracle/shim_ros — TEST INFRASTRUCTURE (see ros/ros.h): the GPS factor is outside the path; declared so that gpsHandler compiles.
#pragma once
namespace GeographicLib { struct LocalCartesian { void Reset(double, double, double) {} void Forward(double, double, double, double& x, double& y, double& z) { x = y = z = 0; } }; }

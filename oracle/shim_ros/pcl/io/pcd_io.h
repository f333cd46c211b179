// oracle/shim_ros — TEST INFRASTRUCTURE (see ros/ros.h): map saving is outside the path, nothing is written.
#pragma once
#include <string>
#include <pcl/point_cloud.h>
namespace pcl { namespace io { template <class T> int savePCDFileBinary(const std::string&, const PointCloud<T>&) { return 0; } } }

// oracle/shim_ros — TEST INFRASTRUCTURE (see ros/ros.h).  PCL point records with PCL's sizes (PointXYZI is 32 bytes: xyz + pad | intensity + pad)
// and the macros the nodes use to declare their own point types.
#pragma once
#include <cstdint>
#include <Eigen/Dense>
#define PCL_ADD_POINT4D float x; float y; float z; float pcl_pad4d_;
#define PCL_ADD_INTENSITY float intensity
#define POINT_CLOUD_REGISTER_POINT_STRUCT(name, fields)
namespace pcl {
struct EIGEN_ALIGN16 PointXYZI { float x = 0, y = 0, z = 0, pad_ = 1.f; float intensity = 0, p1_ = 0, p2_ = 0, p3_ = 0; };
struct EIGEN_ALIGN16 PointXYZINormal { float x = 0, y = 0, z = 0, pad_ = 1.f; float normal_x = 0, normal_y = 0, normal_z = 0, pad2_ = 0; float intensity = 0, curvature = 0, p2_ = 0, p3_ = 0; };
inline float rad2deg(float alpha) { return alpha * 57.29578f; }          // pcl/common/angles.hpp
inline double rad2deg(double alpha) { return alpha * 57.29578; }
}

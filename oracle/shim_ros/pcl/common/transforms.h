// oracle/shim_ros — TEST INFRASTRUCTURE (see ros/ros.h).  pcl::getTransformation / getTranslationAndEulerAngles (pcl/common/eigen.hpp) = the oracle's
// restatement (oracle/liorf_oracle.hpp: get_transformation; roll = atan2(t21, t22), pitch = asin(-t20), yaw = atan2(t10, t00), all float).
#pragma once
#include <cmath>
#include <pcl/point_cloud.h>
namespace pcl {
inline Eigen::Affine3f getTransformation(float x, float y, float z, float roll, float pitch, float yaw) { Eigen::Affine3f t; liorf_oracle::get_transformation(x, y, z, roll, pitch, yaw, t.m); return t; }
inline void getTranslationAndEulerAngles(const Eigen::Affine3f& t, float& x, float& y, float& z, float& roll, float& pitch, float& yaw) {
    x = t(0, 3); y = t(1, 3); z = t(2, 3);
    roll = std::atan2(t(2, 1), t(2, 2)); pitch = std::asin(-t(2, 0)); yaw = std::atan2(t(1, 0), t(0, 0));
}
template <class T> void transformPointCloud(const PointCloud<T>& in, PointCloud<T>& out, const Eigen::Matrix4f& m) {     // publishing only (never reached: no subscribers)
    out = in;
    for (auto& p : out.points) { const float x = p.x, y = p.y, z = p.z; p.x = m(0, 0) * x + m(0, 1) * y + m(0, 2) * z + m(0, 3); p.y = m(1, 0) * x + m(1, 1) * y + m(1, 2) * z + m(1, 3); p.z = m(2, 0) * x + m(2, 1) * y + m(2, 2) * z + m(2, 3); }
}
}

// oracle/shim_ros — TEST INFRASTRUCTURE (see ros/ros.h).  PointCloud2 <-> pcl::PointCloud<T>: the point block is the raw array of T.
#pragma once
#include <cstring>
#include <stdexcept>
#include <pcl/point_cloud.h>
#include <sensor_msgs/PointCloud2.h>
namespace pcl {
template <class T> void toROSMsg(const PointCloud<T>& c, sensor_msgs::PointCloud2& m) {
    m.height = 1; m.width = (uint32_t)c.points.size(); m.point_step = (uint32_t)sizeof(T); m.row_step = m.point_step * m.width; m.is_dense = c.is_dense;
    m.data.resize((size_t)m.row_step);
    if (!c.points.empty()) std::memcpy(m.data.data(), c.points.data(), m.data.size());
}
template <class T> void fromROSMsg(const sensor_msgs::PointCloud2& m, PointCloud<T>& c) {
    if (m.width * m.height != 0 && m.point_step != sizeof(T)) throw std::runtime_error("shim_ros: PointCloud2 point_step differs from the requested point type");
    c.points.resize((size_t)m.width * m.height); c.width = (uint32_t)c.points.size(); c.is_dense = m.is_dense;
    if (!c.points.empty()) std::memcpy((void*)c.points.data(), m.data.data(), c.points.size() * sizeof(T));
}
template <class T> void moveFromROSMsg(sensor_msgs::PointCloud2& m, PointCloud<T>& c) { fromROSMsg(m, c); }
}

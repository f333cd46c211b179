// oracle/shim_ros — TEST INFRASTRUCTURE (see ros/ros.h).  pcl::VoxelGrid<PointXYZI>::filter = the oracle's restatement of PCL's applyFilter
// (oracle/liorf_oracle.hpp: voxel_grid).  Third-party arithmetic: NOT pinned.
#pragma once
#include <pcl/point_cloud.h>
namespace pcl {
template <class T> class VoxelGrid {
public:
    void setLeafSize(float lx, float, float) { leaf_ = lx; }
    void setInputCloud(const typename PointCloud<T>::Ptr& c) { in_ = c; }
    void filter(PointCloud<T>& out) {
        std::vector<liorf_oracle::P4> a(in_->points.size()), o;
        for (size_t i = 0; i < a.size(); ++i) a[i] = liorf_oracle::P4{in_->points[i].x, in_->points[i].y, in_->points[i].z, in_->points[i].intensity};
        liorf_oracle::voxel_grid(a.data(), (int)a.size(), leaf_, o);
        out.points.resize(o.size()); out.width = (uint32_t)o.size();
        for (size_t i = 0; i < o.size(); ++i) { T p; p.x = o[i].x; p.y = o[i].y; p.z = o[i].z; p.intensity = o[i].i; out.points[i] = p; }
    }
private:
    float leaf_ = 1.f; typename PointCloud<T>::Ptr in_;
};
}

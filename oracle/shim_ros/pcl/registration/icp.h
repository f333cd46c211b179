// oracle/shim_ros — TEST INFRASTRUCTURE (see ros/ros.h).  pcl::IterativeClosestPoint: declared so that the loop-closure thread's code compiles; the
// harness never starts that thread (the ICP row has its own oracle, oracle/pyicp.py), align() reports "not converged".
#pragma once
#include <pcl/point_cloud.h>
namespace pcl {
template <class S, class T> class IterativeClosestPoint {
public:
    void setMaxCorrespondenceDistance(double) {} void setMaximumIterations(int) {} void setTransformationEpsilon(double) {} void setEuclideanFitnessEpsilon(double) {}
    void setRANSACIterations(int) {}
    void setInputSource(const typename PointCloud<S>::Ptr&) {} void setInputTarget(const typename PointCloud<T>::Ptr&) {}
    void align(PointCloud<S>&) {}
    bool hasConverged() const { return false; }
    double getFitnessScore() const { return 1e30; }
    Eigen::Matrix4f getFinalTransformation() const { return Eigen::Matrix4f::Identity(); }
};
}

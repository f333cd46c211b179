// oracle/shim_ros — TEST INFRASTRUCTURE (see ros/ros.h).  pcl::PointCloud<T>: the std::vector-like surface the nodes use.
#pragma once
#include <memory>
#include <vector>
#include <pcl/point_types.h>
namespace pcl {
template <class T> struct PointCloud {
    typedef std::shared_ptr<PointCloud<T>> Ptr;
    typedef std::shared_ptr<const PointCloud<T>> ConstPtr;
    std::vector<T> points; uint32_t width = 0, height = 1; bool is_dense = true;
    size_t size() const { return points.size(); }
    bool empty() const { return points.empty(); }
    void clear() { points.clear(); width = 0; }
    void resize(size_t n) { points.resize(n); width = (uint32_t)n; }
    void push_back(const T& p) { points.push_back(p); width = (uint32_t)points.size(); }
    T& back() { return points.back(); }
    const T& back() const { return points.back(); }
    T& front() { return points.front(); }
    const T& front() const { return points.front(); }
    T& operator[](size_t i) { return points[i]; }
    const T& operator[](size_t i) const { return points[i]; }
    typename std::vector<T>::iterator begin() { return points.begin(); }
    typename std::vector<T>::iterator end() { return points.end(); }
    PointCloud& operator+=(const PointCloud& o) { points.insert(points.end(), o.points.begin(), o.points.end()); width = (uint32_t)points.size(); return *this; }
};
template <class A, class B> void copyPointCloud(const PointCloud<A>& in, PointCloud<B>& out) { out.points.assign(in.points.begin(), in.points.end()); out.width = in.width; out.is_dense = in.is_dense; }
namespace console { enum VERBOSITY_LEVEL { L_ALWAYS, L_ERROR, L_WARN, L_INFO, L_DEBUG, L_VERBOSE }; inline void setVerbosityLevel(VERBOSITY_LEVEL) {} }
}

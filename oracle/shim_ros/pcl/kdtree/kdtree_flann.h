// oracle/shim_ros — TEST INFRASTRUCTURE (see ros/ros.h).  pcl::KdTreeFLANN<PointXYZI>: exact k-NN through the reference's own vendored nanoflann
// (KDTreeSingleIndexAdaptor, L2_Simple, leaf 15 — SURVEY §8(d)'s stand-in for FLANN's KDTreeSingleIndex), results ascending by distance (nearest-1: brute force, ties to the lower index);
// radiusSearch = squared L2_Simple distance strictly below r^2, ascending by (distance, index).  Third-party behaviour: NOT pinned.
#pragma once
#include <algorithm>
#include <cmath>
#include <pcl/point_cloud.h>
#include "nanoflann.hpp"
namespace pcl {
template <class T> class KdTreeFLANN {
    struct Adaptor {
        const std::vector<T>* pts = nullptr;
        inline size_t kdtree_get_point_count() const { return pts->size(); }
        inline float kdtree_get_pt(const size_t idx, const size_t dim) const { const T& p = (*pts)[idx]; return dim == 0 ? p.x : (dim == 1 ? p.y : p.z); }
        template <class BBOX> bool kdtree_get_bbox(BBOX&) const { return false; }
    };
    typedef nanoflann::KDTreeSingleIndexAdaptor<nanoflann::L2_Simple_Adaptor<float, Adaptor>, Adaptor, 3, int> Tree;
public:
    typedef std::shared_ptr<KdTreeFLANN<T>> Ptr;
    void setInputCloud(const typename PointCloud<T>::Ptr& c) {
        cloud_ = c; ad_.pts = &cloud_->points; tree_.reset();
        if (!cloud_->points.empty()) { tree_.reset(new Tree(3, ad_, nanoflann::KDTreeSingleIndexAdaptorParams(15))); tree_->buildIndex(); }
    }
    int nearestKSearch(const T& q, int k, std::vector<int>& idx, std::vector<float>& d2) const {
        const int n = (int)cloud_->points.size();
        if (k > n) k = n;                                       // PCL clamps k to the cloud size
        idx.assign(k, 0); d2.assign(k, 0.f);
        if (k == 0) return 0;
        if (k == 1) {   // nearest-1 is only asked on key poses (extractNearby :995, publishGlobalMap :485): the centroid of a two-pose voxel is often EXACTLY
                        // equidistant from both in fp32, and which one FLANN returns is its traversal's business — the canonical rule (distance, index) decides here
            float bd = INFINITY; int best = 0;
            for (int i = 0; i < n; ++i) {
                const T& p = cloud_->points[i];
                float dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z; float d = dx * dx; d += dy * dy; d += dz * dz;
                if (d < bd) { bd = d; best = i; }
            }
            idx[0] = best; d2[0] = bd;
            return 1;
        }
        float qp[3] = {q.x, q.y, q.z};
        nanoflann::KNNResultSet<float, int> rs(k); rs.init(idx.data(), d2.data());
        tree_->findNeighbors(rs, qp, nanoflann::SearchParams());
        return k;
    }
    int radiusSearch(const T& q, double radius, std::vector<int>& idx, std::vector<float>& d2, unsigned = 0) const {
        std::vector<std::pair<float, int>> near;
        const float r2 = (float)(radius * radius);
        for (int i = 0; i < (int)cloud_->points.size(); ++i) {
            const T& p = cloud_->points[i];
            float dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z; float d = dx * dx; d += dy * dy; d += dz * dz;
            if (d < r2) near.emplace_back(d, i);
        }
        std::sort(near.begin(), near.end());
        idx.resize(near.size()); d2.resize(near.size());
        for (size_t i = 0; i < near.size(); ++i) { idx[i] = near[i].second; d2[i] = near[i].first; }
        return (int)near.size();
    }
private:
    typename PointCloud<T>::Ptr cloud_; Adaptor ad_; std::unique_ptr<Tree> tree_;
};
}

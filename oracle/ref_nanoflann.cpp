// ref_nanoflann.cpp — the REFERENCE's own vendored nanoflann 1.3.2 + KDTreeVectorOfVectorsAdaptor, compiled from
// where they lie (/root/reference/include, passed with -I by oracle/Makefile; nothing is copied into this repo).
// Output goes to oracle/_ref/libliorf_ref.so (git-ignored, travels to the GPU box).
//
// TEST INFRASTRUCTURE: pins the oracle's ring-key kNN restatement (include/Scancontext.cpp:270-295) and provides
// the kd-tree based CPU baseline for scan-to-map (PCL's KdTreeFLANN is absent; SURVEY §8(d) names nanoflann's
// KDTreeSingleIndexAdaptor + L2_Simple, leaf 15, as the stand-in).
#include "liorf_oracle.hpp"
#include <memory>
#include <chrono>
#include <cassert>
#include <stdexcept>
#include "nanoflann.hpp"
#include "KDTreeVectorOfVectorsAdaptor.h"
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace liorf_oracle;
using KeyMat = std::vector<std::vector<float>>;                              // include/Scancontext.h:42
using InvKeyTree = KDTreeVectorOfVectorsAdaptor<KeyMat, float>;              // include/Scancontext.h:43

struct Cloud3 {
    const P4* pts; size_t n;
    inline size_t kdtree_get_point_count() const { return n; }
    inline float kdtree_get_pt(const size_t idx, const size_t dim) const { return dim == 0 ? pts[idx].x : (dim == 1 ? pts[idx].y : pts[idx].z); }
    template <class BBOX> bool kdtree_get_bbox(BBOX&) const { return false; }
};
using KD3 = nanoflann::KDTreeSingleIndexAdaptor<nanoflann::L2_Simple_Adaptor<float, Cloud3>, Cloud3, 3, int>;

extern "C" {

// Exactly the call sequence of SCManager::detectLoopClosureID step 1 (include/Scancontext.cpp:270-295):
// tree over keys[0:ntree), leaf 10; KNNResultSet<float>(3) into zero-initialised vectors; SearchParams(10).
void ref_ringkey_knn3(const float* keys, int ntree, const float* q, int nq, int* idx_out, float* d_out) {
    KeyMat mat(ntree, std::vector<float>(20));
    for (int k = 0; k < ntree; ++k) std::memcpy(mat[k].data(), keys + 20 * (size_t)k, 20 * sizeof(float));
    InvKeyTree tree(20, mat, 10);
    for (int i = 0; i < nq; ++i) {
        std::vector<size_t> candidate_indexes(3);
        std::vector<float> out_dists_sqr(3);
        nanoflann::KNNResultSet<float> knnsearch_result(3);
        knnsearch_result.init(&candidate_indexes[0], &out_dists_sqr[0]);
        tree.index->findNeighbors(knnsearch_result, q + 20 * (size_t)i, nanoflann::SearchParams(10));
        for (int j = 0; j < 3; ++j) { idx_out[3 * i + j] = (int)candidate_indexes[j]; d_out[3 * i + j] = out_dists_sqr[j]; }
    }
}

// Tree build only (for timing the every-10th-call rebuild).
double ref_ringkey_tree_build_seconds(const float* keys, int ntree) {
    KeyMat mat(ntree, std::vector<float>(20));
    for (int k = 0; k < ntree; ++k) std::memcpy(mat[k].data(), keys + 20 * (size_t)k, 20 * sizeof(float));
    auto t0 = std::chrono::steady_clock::now();
    InvKeyTree tree(20, mat, 10);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

// CPU baseline of the batched ScanContext search (BASELINE config 5): the reference's own kd-tree (built once per call, as
// detectLoopClosureID rebuilds it every TREE_MAKING_PERIOD_ calls, include/Scancontext.cpp:270-281) over all K ring keys, then per
// query the :289-340 sequence — 3-NN, distanceBtnScanContext of the candidates in kNN order, strict <, threshold — with OpenMP over
// the queries.  timings_s[2] = {tree build, queries}.
void ref_sc_query_batch(const float* keys, const double* descs, int K, const float* qkeys, const double* qdescs, int Q,
                        int* loop_id, int* shift, double* dist, int* cand3, double* timings_s) {
    using clk = std::chrono::steady_clock;
    auto T0 = clk::now();
    KeyMat mat(K, std::vector<float>(20));
    for (int k = 0; k < K; ++k) std::memcpy(mat[k].data(), keys + 20 * (size_t)k, 20 * sizeof(float));
    InvKeyTree tree(20, mat, 10);
    auto T1 = clk::now();
#pragma omp parallel for schedule(dynamic, 4)
    for (int q = 0; q < Q; ++q) {
        std::vector<size_t> ci(3);
        std::vector<float> cd(3);
        nanoflann::KNNResultSet<float> rs(3);
        rs.init(&ci[0], &cd[0]);
        tree.index->findNeighbors(rs, qkeys + 20 * (size_t)q, nanoflann::SearchParams(10));
        double mn = 10000000; int al = 0, nn = 0;
        for (int c = 0; c < 3; ++c) {
            auto r = distance_btn_scancontext(qdescs + 1200 * (size_t)q, descs + 1200 * ci[c]);
            if (r.first < mn) { mn = r.first; al = r.second; nn = (int)ci[c]; }
        }
        loop_id[q] = mn < SC_DIST_THRES ? nn : -1; shift[q] = al; dist[q] = mn;
        if (cand3) for (int c = 0; c < 3; ++c) cand3[3 * q + c] = (int)ci[c];
    }
    if (timings_s) { timings_s[0] = std::chrono::duration<double>(T1 - T0).count(); timings_s[1] = std::chrono::duration<double>(clk::now() - T1).count(); }
}

// 3-D kd-tree 5-NN (stand-in for pcl::KdTreeFLANN::nearestKSearch, src/mapOptmization.cpp:1087,1302)
void ref_kdtree_knn5(const P4* map, int m, const P4* q, int n, int* idx, float* d2) {
    Cloud3 c{map, (size_t)m};
    KD3 tree(3, c, nanoflann::KDTreeSingleIndexAdaptorParams(15));
    tree.buildIndex();
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < n; ++i) {
        float qp[3] = {q[i].x, q[i].y, q[i].z};
        int ii[5] = {-1, -1, -1, -1, -1}; float dd[5] = {INFINITY, INFINITY, INFINITY, INFINITY, INFINITY};
        nanoflann::KNNResultSet<float, int> rs(5); rs.init(ii, dd);
        tree.findNeighbors(rs, qp, nanoflann::SearchParams());
        for (int j = 0; j < 5; ++j) { idx[5 * (size_t)i + j] = ii[j]; d2[5 * (size_t)i + j] = dd[j]; }
    }
}

// CPU baseline for a6: kd-tree rebuilt per call (as :1302 does every frame), OpenMP over points (as :1078),
// serial combine + LM.  timings_s[4] = {kdtree build, surfOptimization total, LM total, whole call}.
int ref_scan2map(const P4* scan, int n, const P4* map, int m, float* tf6, int max_iters, int force_all_iters,
                 float* state37, float* pose_trace, int* nsel_trace, double* timings_s) {
    using clk = std::chrono::steady_clock;
    auto T0 = clk::now();
    if (m < 5 || !(n > 30)) return 0;
    Cloud3 c{map, (size_t)m};
    KD3 tree(3, c, nanoflann::KDTreeSingleIndexAdaptorParams(15));
    tree.buildIndex();
    auto T1 = clk::now();
    double t_surf = 0, t_lm = 0;
    std::vector<P4> coeff(n), ori(n), csel(n); std::vector<uint8_t> flag(n);
    LMState st; st.isDegenerate = state37[0] != 0.f; std::memcpy(st.matP, state37 + 1, 36 * sizeof(float));
    int it = 0;
    for (; it < max_iters; ++it) {
        auto a = clk::now();
        float t[12]; trans2affine(tf6, t);
#pragma omp parallel for schedule(dynamic, 64)
        for (int i = 0; i < n; ++i) {
            P4 sel = apply_affine(t, scan[i]);
            float qp[3] = {sel.x, sel.y, sel.z};
            int ii[5] = {-1, -1, -1, -1, -1}; float dd[5] = {INFINITY, INFINITY, INFINITY, INFINITY, INFINITY};
            nanoflann::KNNResultSet<float, int> rs(5); rs.init(ii, dd);
            tree.findNeighbors(rs, qp, nanoflann::SearchParams());
            P4 cf; flag[i] = surf_point(scan[i], sel, map, ii, dd, cf) ? 1 : 0; coeff[i] = cf;
        }
        auto b = clk::now();
        int nsel = 0;
        for (int i = 0; i < n; ++i) if (flag[i]) { ori[nsel] = scan[i]; csel[nsel] = coeff[i]; ++nsel; }
        bool conv = lm_optimization(it, ori.data(), csel.data(), nsel, tf6, st, nullptr);
        auto d = clk::now();
        t_surf += std::chrono::duration<double>(b - a).count(); t_lm += std::chrono::duration<double>(d - b).count();
        if (pose_trace) std::memcpy(pose_trace + 6 * it, tf6, 6 * sizeof(float));
        if (nsel_trace) nsel_trace[it] = nsel;
        if (conv && !force_all_iters) { ++it; break; }
    }
    state37[0] = st.isDegenerate ? 1.f : 0.f; std::memcpy(state37 + 1, st.matP, 36 * sizeof(float));
    if (timings_s) { timings_s[0] = std::chrono::duration<double>(T1 - T0).count(); timings_s[1] = t_surf; timings_s[2] = t_lm;
                     timings_s[3] = std::chrono::duration<double>(clk::now() - T0).count(); }
    return it;
}

}  // extern "C"

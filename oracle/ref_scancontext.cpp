// ref_scancontext.cpp — extern "C" surface over the REFERENCE's own ScanContext implementation: /root/reference/include/Scancontext.cpp
// is compiled UNCHANGED, from where it lies, next to this file (oracle/Makefile; nothing is copied into the repo) against the header
// stand-ins in oracle/shim/ (Eigen / PCL / OpenCV are not in this image) and the reference's vendored nanoflann.
// Output: oracle/_ref/libliorf_ref_sc.so (git-ignored, travels to the GPU box).
// TEST INFRASTRUCTURE: pins oracle/liorf_oracle.hpp's SCManager restatement (a10-a14) bit for bit — control flow, indexing, thresholds,
// the stale-tree schedule — up to the reduction order of mean / norm / dot, which is the shim's (sequential), not Eigen's.
#include "Scancontext.h"
#include <cstring>
#include <cmath>

extern "C" {

void* refsc_create() { return new SCManager(); }
void refsc_destroy(void* h) { delete (SCManager*)h; }
int refsc_size(void* h) { return (int)((SCManager*)h)->polarcontexts_.size(); }

static pcl::PointCloud<SCPointType> to_cloud(const float* xyzi, int n) {
    pcl::PointCloud<SCPointType> c; c.points.resize((size_t)n);
    for (int i = 0; i < n; ++i) { c.points[i].x = xyzi[4 * i]; c.points[i].y = xyzi[4 * i + 1]; c.points[i].z = xyzi[4 * i + 2]; c.points[i].intensity = xyzi[4 * i + 3]; }
    return c;
}
static void to_rowmajor(const Eigen::MatrixXd& m, double* out) { for (long r = 0; r < m.rows(); ++r) for (long c = 0; c < m.cols(); ++c) out[r * m.cols() + c] = m(r, c); }
static Eigen::MatrixXd from_rowmajor(const double* d, int rows, int cols) { Eigen::MatrixXd m(rows, cols); for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) m(r, c) = d[r * cols + c]; return m; }

// makeScancontext + both keys (include/Scancontext.cpp:151-227); desc row-major [ring][sector]
void refsc_make(void* h, const float* xyzi, int n, double* desc1200, float* ringkey20, double* sectorkey60) {
    SCManager* s = (SCManager*)h;
    auto cloud = to_cloud(xyzi, n);
    Eigen::MatrixXd sc = s->makeScancontext(cloud);
    Eigen::MatrixXd rk = s->makeRingkeyFromScancontext(sc), sk = s->makeSectorkeyFromScancontext(sc);
    to_rowmajor(sc, desc1200);
    std::vector<float> v = eig2stdvec(rk);
    std::memcpy(ringkey20, v.data(), 20 * sizeof(float));
    for (int c = 0; c < 60; ++c) sectorkey60[c] = sk(0, c);
}
void refsc_make_and_save(void* h, const float* xyzi, int n) { auto cloud = to_cloud(xyzi, n); ((SCManager*)h)->makeAndSaveScancontextAndKeys(cloud); }
// appends a ready-made descriptor the way makeAndSaveScancontextAndKeys does (:236-250), keys derived by the reference's functions
void refsc_save_descriptor(void* h, const double* desc1200) {
    SCManager* s = (SCManager*)h;
    Eigen::MatrixXd sc = from_rowmajor(desc1200, 20, 60);
    Eigen::MatrixXd ringkey = s->makeRingkeyFromScancontext(sc), sectorkey = s->makeSectorkeyFromScancontext(sc);
    std::vector<float> vec = eig2stdvec(ringkey);
    s->polarcontexts_.push_back(sc); s->polarcontext_invkeys_.push_back(ringkey); s->polarcontext_vkeys_.push_back(sectorkey); s->polarcontext_invkeys_mat_.push_back(vec);
}
void refsc_get(void* h, int i, double* desc1200, float* key20) {
    SCManager* s = (SCManager*)h;
    if (desc1200) to_rowmajor(s->polarcontexts_[i], desc1200);
    if (key20) std::memcpy(key20, s->polarcontext_invkeys_mat_[i].data(), 20 * sizeof(float));
}
void refsc_detect(void* h, int* loop_id, float* yaw) { auto r = ((SCManager*)h)->detectLoopClosureID(); *loop_id = r.first; *yaw = r.second; }
void refsc_distance(void* h, const double* sc1, const double* sc2, double* dist, int* shift) {
    Eigen::MatrixXd a = from_rowmajor(sc1, 20, 60), b = from_rowmajor(sc2, 20, 60);
    auto r = ((SCManager*)h)->distanceBtnScanContext(a, b); *dist = r.first; *shift = r.second;
}
float refsc_xy2theta(float x, float y) { return xy2theta(x, y); }

// A BATCH of queries against the stored database through the reference's own detectLoopClosureID (:253-344), unchanged, one call per query (bench.py --impl
// reference at N > 1).  The function answers for the LAST stored entry against a tree over entries [0, size - 30) (:270-281), so the harness keeps 29 padding
// entries behind the K database rows, appends the query the way makeAndSaveScancontextAndKeys does (:236-250), calls the function and removes the query again;
// the tree is built by the function itself on the first call (tree_making_period_conter == 0) and kept afterwards (the counter is held off the rebuild period:
// the database does not change between queries).  shift = yaw / 6 degrees.  Single-threaded, as the reference's loop-closure thread is.
void refsc_query_batch_own(void* h, const double* qdescs, int Q, int* loop_id, int* shift) {
    SCManager* s = (SCManager*)h;
    const size_t K = s->polarcontexts_.size();
    for (int i = 0; i < 29; ++i) {                                  // padding: never inside the tree (the last 30 entries are excluded), never a candidate
        s->polarcontexts_.push_back(s->polarcontexts_[0]); s->polarcontext_invkeys_.push_back(s->polarcontext_invkeys_[0]);
        s->polarcontext_vkeys_.push_back(s->polarcontext_vkeys_[0]); s->polarcontext_invkeys_mat_.push_back(s->polarcontext_invkeys_mat_[0]);
    }
    s->tree_making_period_conter = 0;
    for (int q = 0; q < Q; ++q) {
        Eigen::MatrixXd sc = from_rowmajor(qdescs + (size_t)q * 1200, 20, 60);
        Eigen::MatrixXd ringkey = s->makeRingkeyFromScancontext(sc), sectorkey = s->makeSectorkeyFromScancontext(sc);
        std::vector<float> vec = eig2stdvec(ringkey);
        s->polarcontexts_.push_back(sc); s->polarcontext_invkeys_.push_back(ringkey); s->polarcontext_vkeys_.push_back(sectorkey); s->polarcontext_invkeys_mat_.push_back(vec);
        std::pair<int, float> r = s->detectLoopClosureID();
        if (s->tree_making_period_conter % 10 == 0) s->tree_making_period_conter = 1;
        loop_id[q] = r.first; shift[q] = (int)std::lround((double)r.second * 180.0 / M_PI / 6.0);
        s->polarcontexts_.pop_back(); s->polarcontext_invkeys_.pop_back(); s->polarcontext_vkeys_.pop_back(); s->polarcontext_invkeys_mat_.pop_back();
    }
    s->polarcontexts_.resize(K); s->polarcontext_invkeys_.resize(K); s->polarcontext_vkeys_.resize(K); s->polarcontext_invkeys_mat_.resize(K);
}

}  // extern "C"

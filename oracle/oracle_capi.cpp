// oracle_capi.cpp — extern "C" surface of the CPU oracle for ctypes (tests / smoke / bench cpu_baseline only).
// TEST INFRASTRUCTURE — never linked into the product library.
#include "liorf_oracle.hpp"
#include <chrono>
#include <cstring>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace liorf_oracle;

extern "C" {

int orc_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

// scalar pre/post steps of laserCloudInfoHandler: state37 = {tf[6], lastImuTransformation[12], lastImuPreTransformation[12], available}
static void st_load(MapOptScalarState& m, const float* st) { std::memcpy(m.transformTobeMapped, st, 24); std::memcpy(m.lastImuTransformation, st + 6, 48); std::memcpy(m.lastImuPreTransformation, st + 18, 48); m.lastImuPreTransAvailable = (int)st[30]; }
static void st_store(const MapOptScalarState& m, float* st) { std::memcpy(st, m.transformTobeMapped, 24); std::memcpy(st + 6, m.lastImuTransformation, 48); std::memcpy(st + 18, m.lastImuPreTransformation, 48); st[30] = (float)m.lastImuPreTransAvailable; }
void orc_update_initial_guess(float* state31, int no_keyframes, const float* ci11 /*imuAvail, odomAvail, imu rpy, guess xyz rpy*/, int useImuHeading, int imuType) {
    MapOptScalarState m; st_load(m, state31);
    GuessCloudInfo ci{(int)ci11[0], (int)ci11[1], ci11[2], ci11[3], ci11[4], ci11[5], ci11[6], ci11[7], ci11[8], ci11[9], ci11[10]};
    m.updateInitialGuess(no_keyframes != 0, ci, useImuHeading != 0, imuType);
    st_store(m, state31);
}
void orc_transform_update(float* tf6, int imuAvailable, int imuType, float imuRoll, float imuPitch, float weight, float rot_tol, float z_tol) {
    MapOptScalarState m; std::memcpy(m.transformTobeMapped, tf6, 24);
    GuessCloudInfo ci{imuAvailable, 0, imuRoll, imuPitch, 0, 0, 0, 0, 0, 0, 0};
    m.transformUpdate(ci, imuType, weight, rot_tol, z_tol);
    std::memcpy(tf6, m.transformTobeMapped, 24);
}

void orc_get_transformation(float x, float y, float z, float roll, float pitch, float yaw, float* t12) {
    get_transformation(x, y, z, roll, pitch, yaw, t12);
}

// pose6 = (roll,pitch,yaw,x,y,z) like transformTobeMapped / PointTypePose order used by trans2Affine3f
void orc_transform_cloud(const P4* in, int n, const float* pose6, P4* out) {
    float t[12]; trans2affine(pose6, t);
#pragma omp parallel for
    for (int i = 0; i < n; ++i) out[i] = apply_affine(t, in[i]);
}

// returns n_out (or -1: PCL overflow guard → caller must treat output == input)
int orc_voxel_grid(const P4* in, int n, float leaf, P4* out, int* membership, int* out_keys, int* meta_minb_divb /*6 or null*/) {
    std::vector<P4> o; std::vector<int> mem, keys; VoxelMeta vm;
    int r = voxel_grid(in, n, leaf, o, membership ? &mem : nullptr, out_keys ? &keys : nullptr, &vm);
    if (r < 0) return r;
    if (out) std::memcpy(out, o.data(), o.size() * sizeof(P4));
    if (membership) std::memcpy(membership, mem.data(), mem.size() * sizeof(int));
    if (out_keys) std::memcpy(out_keys, keys.data(), keys.size() * sizeof(int));
    if (meta_minb_divb && n > 0) for (int a = 0; a < 3; ++a) { meta_minb_divb[a] = vm.min_b[a]; meta_minb_divb[3 + a] = vm.div_b[a]; }
    return r;
}

void orc_knn5(const P4* map, int m, const P4* q, int n, int* idx, float* d2) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < n; ++i) knn5_brute(map, m, q[i], idx + 5 * (size_t)i, d2 + 5 * (size_t)i);
}

void orc_colpiv_qr_solve_5x3(const float* A, const float* b, float* x) { colpiv_qr_solve_5x3(A, b, x); }
int orc_cv_qr_solve6(const float* A, const float* b, float* x) { return cv_qr_solve6(A, b, x) ? 1 : 0; }
void orc_cv_jacobi6(const float* A, float* W, float* V) { cv_jacobi6(A, W, V); }
int orc_cv_lu_invert6(const float* A, float* inv) { return cv_lu_invert6(A, inv) ? 1 : 0; }
void orc_cv_gemm6(const float* A, const float* B, float* C) { cv_gemm6(A, B, C); }

// a7 over the whole scan.  Outputs are per input point i (like laserCloudOriSurfVec/coeffSelSurfVec/Flag).
// idx/d2/plane/pointSel are optional (nullable) diagnostics.
void orc_surf_optimization(const P4* scan, int n, const P4* map, int m, const float* tf6, P4* coeff, uint8_t* flag,
                           int* idx_out, float* d2_out, float* plane_out, P4* sel_out) {
    float t[12]; trans2affine(tf6, t);
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < n; ++i) {
        P4 sel = apply_affine(t, scan[i]);
        int idx[5]; float d2[5]; knn5_brute(map, m, sel, idx, d2);
        P4 c; float pl[4] = {0, 0, 0, 0};
        bool f = m >= 5 && surf_point(scan[i], sel, map, idx, d2, c, pl);
        if (!f) { /* reference leaves stale entries; flag false means ignored */ }
        coeff[i] = c; flag[i] = f ? 1 : 0;
        if (idx_out) for (int j = 0; j < 5; ++j) idx_out[5 * (size_t)i + j] = idx[j];
        if (d2_out) for (int j = 0; j < 5; ++j) d2_out[5 * (size_t)i + j] = d2[j];
        if (plane_out) for (int j = 0; j < 4; ++j) plane_out[4 * (size_t)i + j] = pl[j];
        if (sel_out) sel_out[i] = sel;
    }
}

// a8: ascending-i compaction
int orc_combine_optimization_coeffs(const P4* scan, const P4* coeff, const uint8_t* flag, int n, P4* ori_out, P4* coeff_out) {
    int k = 0;
    for (int i = 0; i < n; ++i) if (flag[i]) { ori_out[k] = scan[i]; coeff_out[k] = coeff[i]; ++k; }
    return k;
}

// a9.  state = [isDegenerate (as float 0/1), matP[36]] (37 floats, in/out).  trace: AtA[36] AtB[6] X[6] nsel conv degen solved (52 floats)
int orc_lm_optimization(int iter, const P4* ori, const P4* coeff, int nsel, float* tf6, float* state37, float* trace52) {
    LMState st; st.isDegenerate = state37[0] != 0.f; std::memcpy(st.matP, state37 + 1, 36 * sizeof(float));
    LMTrace tr; bool c = lm_optimization(iter, ori, coeff, nsel, tf6, st, &tr);
    state37[0] = st.isDegenerate ? 1.f : 0.f; std::memcpy(state37 + 1, st.matP, 36 * sizeof(float));
    if (trace52) {
        std::memcpy(trace52, tr.AtA, 36 * sizeof(float)); std::memcpy(trace52 + 36, tr.AtB, 6 * sizeof(float)); std::memcpy(trace52 + 42, tr.X, 6 * sizeof(float));
        trace52[48] = (float)tr.nsel; trace52[49] = (float)tr.converged; trace52[50] = (float)tr.degenerate; trace52[51] = (float)tr.solved;
    }
    return c ? 1 : 0;
}

// a6: scan2MapOptimization loop (:1295-1321) with brute-force exact 5-NN.  pose_trace: (max_iters x 6) pose after each
// iteration; nsel_trace: max_iters ints.  Returns number of iterations executed.  Guard n > 30 per :1300.
int orc_scan2map(const P4* scan, int n, const P4* map, int m, float* tf6, int max_iters, int force_all_iters,
                 float* state37, float* pose_trace, int* nsel_trace) {
    if (m < 1 || !(n > 30)) return 0;
    std::vector<P4> coeff(n), ori(n), csel(n); std::vector<uint8_t> flag(n);
    int it = 0;
    for (; it < max_iters; ++it) {
        orc_surf_optimization(scan, n, map, m, tf6, coeff.data(), flag.data(), nullptr, nullptr, nullptr, nullptr);
        int nsel = orc_combine_optimization_coeffs(scan, coeff.data(), flag.data(), n, ori.data(), csel.data());
        int conv = orc_lm_optimization(it, ori.data(), csel.data(), nsel, tf6, state37, nullptr);
        if (pose_trace) std::memcpy(pose_trace + 6 * it, tf6, 6 * sizeof(float));
        if (nsel_trace) nsel_trace[it] = nsel;
        if (conv && !force_all_iters) { ++it; break; }
    }
    return it;
}

int orc_project_point_cloud(const PRaw* in, int n, float minR, float maxR, int N_SCAN, int dsRate, int pfn, double timeScanCur,
                            const double* imuTime, const double* rx, const double* ry, const double* rz, int imuPointerCur,
                            int deskew_enabled, P4* out, int* kept_index) {
    DeskewParams P{minR, maxR, N_SCAN, dsRate, pfn};
    std::vector<P4> o; std::vector<int> ki;
    int r = project_point_cloud(in, n, P, timeScanCur, imuTime, rx, ry, rz, imuPointerCur, deskew_enabled, o, &ki);
    if (out) std::memcpy(out, o.data(), o.size() * sizeof(P4));
    if (kept_index) std::memcpy(kept_index, ki.data(), ki.size() * sizeof(int));
    return r;
}

// imuDeskewInfo over a deque (returns imuAvailable; rows = imuPointerCur + 1; queue_left = samples still queued afterwards)
int orc_imu_deskew_info(const double* stamp, const double* gyro_xyz, int n, double timeScanCur, double timeScanEnd, double* imuTime, double* rx, double* ry, double* rz,
                        int* imuPointerCur, int* queue_left) {
    std::deque<ImuSample> q;
    for (int i = 0; i < n; ++i) q.push_back(ImuSample{stamp[i], gyro_xyz[3 * i], gyro_xyz[3 * i + 1], gyro_xyz[3 * i + 2]});
    std::vector<double> t(2000), x(2000), y(2000), z(2000);
    int ptr = -1;
    bool ok = imu_deskew_info(q, timeScanCur, timeScanEnd, t, x, y, z, ptr);
    for (int i = 0; i <= ptr && i < 2000; ++i) { imuTime[i] = t[i]; rx[i] = x[i]; ry[i] = y[i]; rz[i] = z[i]; }
    *imuPointerCur = ptr; *queue_left = (int)q.size();
    return ok ? 1 : 0;
}

// ---- ScanContext ----
void orc_sc_make(const P4* pts, int n, double* desc1200, float* ringkey20, double* sectorkey60) {
    make_scancontext(pts, n, desc1200);
    double rk[SC_RING]; make_ringkey(desc1200, rk);
    if (ringkey20) for (int r = 0; r < SC_RING; ++r) ringkey20[r] = (float)rk[r];
    if (sectorkey60) make_sectorkey(desc1200, sectorkey60);
}
void orc_sc_keys_from_desc(const double* desc1200, float* ringkey20, double* sectorkey60) {
    double rk[SC_RING]; make_ringkey(desc1200, rk);
    if (ringkey20) for (int r = 0; r < SC_RING; ++r) ringkey20[r] = (float)rk[r];
    if (sectorkey60) make_sectorkey(desc1200, sectorkey60);
}
void orc_sc_distance(const double* sc1, const double* sc2, double* dist, int* shift) {
    auto r = distance_btn_scancontext(sc1, sc2); *dist = r.first; *shift = r.second;
}
void orc_ringkey_top3(const float* keys, int ntree, const float* q, int nq, int* idx, float* d) {
#pragma omp parallel for schedule(dynamic, 8)
    for (int i = 0; i < nq; ++i) ringkey_top3(keys, ntree, q + 20 * (size_t)i, idx + 3 * (size_t)i, d + 3 * (size_t)i);
}
// batched query of config 5: for each query: exact top-3 ring-key candidates over the K-entry DB (by (dist, idx)),
// then distanceBtnScanContext for the candidates in kNN order with strict-< argmin, threshold SC_DIST_THRES.
void orc_sc_query_batch(const float* keys, const double* descs, int K, const float* qkeys, const double* qdescs, int Q,
                        int* loop_id, int* shift, double* dist, int* cand3) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int q = 0; q < Q; ++q) {
        int ci[3]; float cd[3]; ringkey_top3(keys, K, qkeys + 20 * (size_t)q, ci, cd);
        double mn = 10000000; int al = 0, nn = 0;
        for (int c = 0; c < 3; ++c) {
            auto r = distance_btn_scancontext(qdescs + 1200 * (size_t)q, descs + 1200 * (size_t)ci[c]);
            if (r.first < mn) { mn = r.first; al = r.second; nn = ci[c]; }
        }
        loop_id[q] = mn < SC_DIST_THRES ? nn : -1; shift[q] = al; dist[q] = mn;
        if (cand3) { cand3[3 * q] = ci[0]; cand3[3 * q + 1] = ci[1]; cand3[3 * q + 2] = ci[2]; }
    }
}
void* orc_sc_create() { return new SCManager(); }
void orc_sc_destroy(void* h) { delete (SCManager*)h; }
void orc_sc_make_and_save(void* h, const P4* pts, int n) { ((SCManager*)h)->makeAndSaveScancontextAndKeys(pts, n); }
void orc_sc_save_descriptor(void* h, const double* desc) { ((SCManager*)h)->saveDescriptor(desc); }
int orc_sc_size(void* h) { return (int)((SCManager*)h)->polarcontexts_.size(); }
void orc_sc_get(void* h, int i, double* desc1200, float* key20) {
    SCManager* s = (SCManager*)h;
    if (desc1200) std::memcpy(desc1200, s->polarcontexts_[i].data(), 1200 * sizeof(double));
    if (key20) std::memcpy(key20, s->invkeys_mat_[i].data(), 20 * sizeof(float));
}
void orc_sc_detect(void* h, int* loop_id, float* yaw, double* min_dist, int* cand3) {
    auto r = ((SCManager*)h)->detectLoopClosureID(min_dist, cand3); *loop_id = r.first; *yaw = r.second;
}

}  // extern "C"

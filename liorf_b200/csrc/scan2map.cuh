// scan2map.cuh — scan-to-map registration kernels.
//   surfOptimization          src/mapOptmization.cpp:1074-1143   (exact 5-NN on the voxel-hash grid + 5x3 plane fit + weights)
//   combineOptimizationCoeffs src/mapOptmization.cpp:1145-1156   (order-preserving compaction — hook only; fused away in the solver)
//   LMOptimization            src/mapOptmization.cpp:1158-1293   (Jacobian rows → 6x6 normal equations → QR solve → degeneracy projector)
//   scan2MapOptimization      src/mapOptmization.cpp:1295-1321   (≤30 iterations, ONE persistent cooperative kernel, no host round trips)
//
// Thread mapping: LIORF_G (= 8) lanes cooperate on one query point.  The group strides over the candidate points of the
// 27 neighbouring cells (9 contiguous x-runs), keeps per-lane top-5 lists ordered by (distance, original index), merges
// them with sub-warp shuffles, then every lane of the group redundantly evaluates the plane fit (latency, not ALU, is
// the bound) and the 28 normal-equation products are split 4-per-lane across the group and accumulated in fp64.
#pragma once
#include <cooperative_groups.h>
#include "linalg.cuh"
#include "localmap.cuh"

namespace liorf {
namespace cg = cooperative_groups;

constexpr int LIORF_G = 8;                 // lanes per query
constexpr int S2M_BLOCK = 256;             // per-function hook kernels
constexpr int S2M_QPB = S2M_BLOCK / LIORF_G;
constexpr int S2M_MAX_ITERS = 64;
constexpr int NPROD = 28;                  // 21 upper-triangular AtA + 6 AtB + 1 count

struct LMDeviceState {                     // persists across iterations AND frames (members :139-140)
    int isDegenerate;
    float matP[36];
};
struct S2MTrace {                          // mirrors liorf_lm_trace in include/liorf_b200.h
    float pose[S2M_MAX_ITERS][6];
    int nsel[S2M_MAX_ITERS];
    int iters, converged, degenerate, ran;
};

struct Top5 { float d[5]; int oi[5]; int pos[5]; };

__device__ __forceinline__ void top5_init(Top5& t) {
#pragma unroll
    for (int j = 0; j < 5; ++j) { t.d[j] = INFINITY; t.oi[j] = 0x7fffffff; t.pos[j] = -1; }
}
__device__ __forceinline__ bool less_di(float d, int i, float d2, int i2) { return d < d2 || (d == d2 && i < i2); }
__device__ __forceinline__ void top5_insert(Top5& t, float d, int oi, int pos) {
    if (!less_di(d, oi, t.d[4], t.oi[4])) return;
    t.d[4] = d; t.oi[4] = oi; t.pos[4] = pos;
#pragma unroll
    for (int j = 4; j > 0; --j) {
        if (less_di(t.d[j], t.oi[j], t.d[j - 1], t.oi[j - 1])) {
            float td = t.d[j]; t.d[j] = t.d[j - 1]; t.d[j - 1] = td;
            int ti = t.oi[j]; t.oi[j] = t.oi[j - 1]; t.oi[j - 1] = ti;
            int tp = t.pos[j]; t.pos[j] = t.pos[j - 1]; t.pos[j - 1] = tp;
        }
    }
}

__device__ __forceinline__ void knn_scan_range(const float4* __restrict__ gmap, unsigned b, unsigned e, int gl, const float4& q, Top5& t) {
    for (unsigned k = b + gl; k < e; k += LIORF_G) {
        float4 p = __ldg(gmap + k);
        float dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z;
        float d = dx * dx; d += dy * dy; d += dz * dz;              // FLANN L2_Simple op order, no FMA
        if (d < 1.0f) top5_insert(t, d, __float_as_int(p.w), (int)k);
    }
}

// Exact 5 nearest map points with squared distance < 1.0, ordered by (distance, original index).  All lanes of the
// group return identical lists.  Slots beyond the number found hold d = +inf, oi = INT_MAX, pos = -1.
// Latency structure: the 18 cell-range bounds of the 9 x-runs are fetched together (one L2 round trip), then the
// concatenated candidate list is walked KNN_U candidates per lane per step with the loads issued before any compare.
constexpr int KNN_U = 4;
__device__ __forceinline__ void knn5_group(const float4 q, const unsigned* __restrict__ cell_start, const float4* __restrict__ gmap,
                                           GridDims g, Top5& res) {
    const int gl = threadIdx.x & (LIORF_G - 1);
    Top5 mine; top5_init(mine);
    const int cx = (int)floorf(q.x), cy = (int)floorf(q.y), cz = (int)floorf(q.z);
    const int x0 = (cx - 1) & (g.DX - 1), x1 = cx & (g.DX - 1), x2 = (cx + 1) & (g.DX - 1);
    const bool contiguous = (x0 + 2 == x2);
    if (contiguous) {
        unsigned bs[9], P[10];
        P[0] = 0;
#pragma unroll
        for (int r = 0; r < 9; ++r) {
            const int dz = r / 3 - 1, dy = r % 3 - 1;
            const int row = (((cz + dz) & (g.DZ - 1)) * g.DY + ((cy + dy) & (g.DY - 1))) * g.DX;
            bs[r] = __ldg(cell_start + row + x0);
            P[r + 1] = __ldg(cell_start + row + x2 + 1);          // end of the run, turned into a prefix below
        }
#pragma unroll
        for (int r = 0; r < 9; ++r) P[r + 1] = P[r] + (P[r + 1] - bs[r]);
        const unsigned T = P[9];
        for (unsigned f0 = gl; f0 < T; f0 += LIORF_G * KNN_U) {
            float4 p[KNN_U]; unsigned pos[KNN_U];
#pragma unroll
            for (int u = 0; u < KNN_U; ++u) {
                const unsigned f = f0 + LIORF_G * u;
                unsigned base = bs[0], pb = 0;
#pragma unroll
                for (int rr = 1; rr < 9; ++rr) if (f >= P[rr]) { base = bs[rr]; pb = P[rr]; }
                pos[u] = base + (f - pb);
                p[u] = f < T ? __ldg(gmap + pos[u]) : make_float4(1e30f, 1e30f, 1e30f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < KNN_U; ++u) {
                float dx = q.x - p[u].x, dy = q.y - p[u].y, dz = q.z - p[u].z;
                float d = dx * dx; d += dy * dy; d += dz * dz;      // FLANN L2_Simple op order, no FMA
                if (d < 1.0f) top5_insert(mine, d, __float_as_int(p[u].w), (int)pos[u]);
            }
        }
    } else {                                                        // x-run wraps around the torus: three separate cells per row
#pragma unroll 1
        for (int r = 0; r < 9; ++r) {
            const int dz = r / 3 - 1, dy = r % 3 - 1;
            const int row = (((cz + dz) & (g.DZ - 1)) * g.DY + ((cy + dy) & (g.DY - 1))) * g.DX;
            knn_scan_range(gmap, __ldg(cell_start + row + x0), __ldg(cell_start + row + x0 + 1), gl, q, mine);
            knn_scan_range(gmap, __ldg(cell_start + row + x1), __ldg(cell_start + row + x1 + 1), gl, q, mine);
            knn_scan_range(gmap, __ldg(cell_start + row + x2), __ldg(cell_start + row + x2 + 1), gl, q, mine);
        }
    }
    // merge the LIORF_G sorted lists: 5 rounds of group arg-min by (d, oi); the winning lane pops its head
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        float md = mine.d[0]; int mi = mine.oi[0], mp = mine.pos[0];
#pragma unroll
        for (int o = LIORF_G / 2; o > 0; o >>= 1) {
            float od = __shfl_xor_sync(FULL, md, o); int oi = __shfl_xor_sync(FULL, mi, o); int op = __shfl_xor_sync(FULL, mp, o);
            if (less_di(od, oi, md, mi)) { md = od; mi = oi; mp = op; }
        }
        if (mine.d[0] == md && mine.oi[0] == mi) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { mine.d[j] = mine.d[j + 1]; mine.oi[j] = mine.oi[j + 1]; mine.pos[j] = mine.pos[j + 1]; }
            mine.d[4] = INFINITY; mine.oi[4] = 0x7fffffff; mine.pos[4] = -1;
        }
        res.d[r] = md; res.oi[r] = mi; res.pos[r] = mp;
    }
}

// surfOptimization body after the neighbour search (:1089-1139).  Returns the flag; coeff = (s*pa, s*pb, s*pc, s*pd2).
__device__ __forceinline__ bool surf_point_dev(const float4 pointOri, const float4 pointSel, const float4* __restrict__ gmap, const Top5& nn,
                                               float4& coeff, float* plane /*4 or null*/) {
    coeff = make_float4(0.f, 0.f, 0.f, 0.f);
    if (plane) { plane[0] = plane[1] = plane[2] = plane[3] = 0.f; }
    if (nn.pos[4] < 0 || !((double)nn.d[4] < 1.0)) return false;                 // :1097
    float A[5][3];
#pragma unroll
    for (int j = 0; j < 5; ++j) { float4 m = __ldg(gmap + nn.pos[j]); A[j][0] = m.x; A[j][1] = m.y; A[j][2] = m.z; }
    float x[3];
    colpiv_qr_solve_5x3(A, x);                                                   // :1104
    float pa = x[0], pb = x[1], pc = x[2], pd = 1.f;
    float ps = sqrtf(pa * pa + pb * pb + pc * pc);                               // :1111
    pa /= ps; pb /= ps; pc /= ps; pd /= ps;
    if (plane) { plane[0] = pa; plane[1] = pb; plane[2] = pc; plane[3] = pd; }
    bool planeValid = true;
#pragma unroll
    for (int j = 0; j < 5; ++j)                                                  // :1115-1122 (float → double compare with 0.2)
        if ((double)fabsf(pa * A[j][0] + pb * A[j][1] + pc * A[j][2] + pd) > 0.2) planeValid = false;
    if (!planeValid) return false;
    float pd2 = pa * pointSel.x + pb * pointSel.y + pc * pointSel.z + pd;        // :1125
    float rr = sqrtf(sqrtf(pointOri.x * pointOri.x + pointOri.y * pointOri.y + pointOri.z * pointOri.z));
    float s = (float)(1.0 - 0.9 * (double)fabsf(pd2) / (double)rr);              // :1127-1128 (double expression)
    coeff = make_float4(s * pa, s * pb, s * pc, s * pd2);                        // :1130-1133
    return (double)s > 0.1;                                                      // :1135
}

struct LMTrig { float srx, crx, sry, cry, srz, crz; };
__device__ __forceinline__ LMTrig lm_trig(const float* tf) {                     // :1170-1175
    LMTrig t; t.srx = sin_f(tf[2]); t.crx = cos_f(tf[2]); t.sry = sin_f(tf[1]); t.cry = cos_f(tf[1]); t.srz = sin_f(tf[0]); t.crz = cos_f(tf[0]);
    return t;
}
// Jacobian row (:1216-1234): row = (arz, ary, arx, cx, cy, cz), b = -coeff.intensity
__device__ __forceinline__ void lm_row_dev(const LMTrig& g, const float4 p, const float4 c, float row[7]) {
    const float srx = g.srx, crx = g.crx, sry = g.sry, cry = g.cry, srz = g.srz, crz = g.crz;
    float arx = (-srx * cry * p.x - (srx * sry * srz + crx * crz) * p.y + (crx * srz - srx * sry * crz) * p.z) * c.x
              + (crx * cry * p.x - (srx * crz - crx * sry * srz) * p.y + (crx * sry * crz + srx * srz) * p.z) * c.y;
    float ary = (-crx * sry * p.x + crx * cry * srz * p.y + crx * cry * crz * p.z) * c.x
              + (-srx * sry * p.x + srx * sry * srz * p.y + srx * cry * crz * p.z) * c.y
              + (-cry * p.x - sry * srz * p.y - sry * crz * p.z) * c.z;
    float arz = ((crx * sry * crz + srx * srz) * p.y + (srx * crz - crx * sry * srz) * p.z) * c.x
              + ((-crx * srz + srx * sry * crz) * p.y + (-srx * sry * srz - crx * crz) * p.z) * c.y
              + (cry * crz * p.y - cry * srz * p.z) * c.z;
    row[0] = arz; row[1] = ary; row[2] = arx; row[3] = c.x; row[4] = c.y; row[5] = c.z; row[6] = -c.w;
}

// product p of the 28: (i, j) indices into the 8-vector v = (row[0..5], b, 1):
// p 0..20 = upper triangle of AtA row by row, 21..26 = AtB, 27 = 1*1 (count of selected rows)
__host__ __device__ constexpr int prod_i(int p) { return p < 6 ? 0 : p < 11 ? 1 : p < 15 ? 2 : p < 18 ? 3 : p < 20 ? 4 : p < 21 ? 5 : p < 27 ? p - 21 : 7; }
__host__ __device__ constexpr int prod_j(int p) { return p < 6 ? p : p < 11 ? p - 5 : p < 15 ? p - 9 : p < 18 ? p - 12 : p < 20 ? p - 14 : p < 21 ? 5 : p < 27 ? 6 : 7; }

// iteration-0 only (:1242-1264): cv::eigen → zero the eigenvector rows of eigenvalues < 100 → matP = V^-1 * V2
__device__ __noinline__ void lm_degeneracy_dev(const float* AtA, LMDeviceState* st, float* scratchA /*36*/, float* scratchV /*36*/) {
    float W[6], V2[36], Vinv[36];
    for (int i = 0; i < 36; ++i) scratchA[i] = AtA[i];
    jacobi6(scratchA, W, scratchV);
    for (int i = 0; i < 36; ++i) V2[i] = scratchV[i];
    int deg = 0;
    for (int i = 5; i >= 0; --i) {
        if (W[i] < 100.f) { for (int j = 0; j < 6; ++j) V2[i * 6 + j] = 0.f; deg = 1; }
        else break;
    }
    for (int i = 0; i < 36; ++i) scratchA[i] = scratchV[i];
    lu_invert6(scratchA, Vinv);
    gemm6(Vinv, V2, st->matP);
    st->isDegenerate = deg;
}

// Certificate of non-degeneracy: true ⇒ every eigenvalue of the symmetric fp32 matrix AtA exceeds mu, proven by an fp64
// LDL^T factorisation of (AtA - mu I) with strictly positive pivots.  mu = 100 (the reference's eigen threshold, :1252)
// plus a margin that dominates the fp32 Jacobi solver's eigenvalue error (<= ~n eps ||A||), so cv::eigen would also
// report all eigenvalues >= 100 and isDegenerate = false.  When the certificate fails the full Jacobi path runs.
__device__ __forceinline__ bool certify_non_degenerate(const float* AtA) {
    double trace = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) trace += (double)AtA[i * 7];
    const double mu = 100.0 + 2e-5 * trace;
    double L[6][6], D[6];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double d = (double)AtA[j * 7] - mu;
#pragma unroll
        for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k] * D[k];
        D[j] = d;
        if (!(d > 1e-9 * trace)) ok = false;
#pragma unroll
        for (int i = j + 1; i < 6; ++i) {
            double v = (double)AtA[i * 6 + j];
#pragma unroll
            for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k] * D[k];
            L[i][j] = v / d;
        }
    }
    return ok;
}

// Solve step shared by the hook kernel and the persistent kernel; executed by ONE thread, everything in registers
// except the iteration-0 eigen step.  sums: 28 fp64 totals.  Updates tf (6), state; returns converged.
// FAST_DEGENERACY: use the certificate above to skip cv::eigen when the system is provably well conditioned
// (matP is then set to the identity; it is only ever read when isDegenerate is true).
template <bool FAST_DEGENERACY>
__device__ __forceinline__ bool lm_solve_dev(int iter, const double* sums, float* tf, LMDeviceState* st, float* scratchA /*36*/, float* scratchV /*36*/,
                                             float* AtA_out, float* AtB_out, float* X_out, int* nsel_out) {
    float AtA[36], AtB[6], X[6];
#pragma unroll
    for (int p = 0; p < 21; ++p) { const float v = (float)sums[p]; AtA[prod_i(p) * 6 + prod_j(p)] = v; AtA[prod_j(p) * 6 + prod_i(p)] = v; }
#pragma unroll
    for (int i = 0; i < 6; ++i) AtB[i] = (float)sums[21 + i];
    const int nsel = (int)(sums[27] + 0.5);
    if (nsel_out) *nsel_out = nsel;
    if (AtA_out) {
#pragma unroll
        for (int i = 0; i < 36; ++i) AtA_out[i] = AtA[i];
    }
    if (AtB_out) {
#pragma unroll
        for (int i = 0; i < 6; ++i) AtB_out[i] = AtB[i];
    }
    if (X_out) {
#pragma unroll
        for (int i = 0; i < 6; ++i) X_out[i] = 0.f;
    }
    if (nsel < 50) return false;                                                    // :1178
    qr_solve6(AtA, AtB, X);                                                         // :1240
    if (iter == 0) {                                                                // :1242-1264
        if (FAST_DEGENERACY && certify_non_degenerate(AtA)) {
            st->isDegenerate = 0;
#pragma unroll
            for (int i = 0; i < 36; ++i) st->matP[i] = (i % 7 == 0) ? 1.f : 0.f;
        } else {
            float tmpA[36];
#pragma unroll
            for (int i = 0; i < 36; ++i) tmpA[i] = AtA[i];
            lm_degeneracy_dev(tmpA, st, scratchA, scratchV);
        }
    }
    if (st->isDegenerate) {                                                         // :1266-1271
        float X2[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) X2[i] = X[i];
        gemv6(st->matP, X2, X);
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) tf[i] += X[i];                                      // :1273-1278
    if (X_out) {
#pragma unroll
        for (int i = 0; i < 6; ++i) X_out[i] = X[i];
    }
    const float r2d = 57.29578f;
    double a0 = (double)(X[0] * r2d), a1 = (double)(X[1] * r2d), a2 = (double)(X[2] * r2d);
    double t0 = (double)(X[3] * 100), t1 = (double)(X[4] * 100), t2 = (double)(X[5] * 100);
    float deltaR = (float)sqrt(a0 * a0 + a1 * a1 + a2 * a2);                        // :1280-1287
    float deltaT = (float)sqrt(t0 * t0 + t1 * t1 + t2 * t2);
    return (double)deltaR < 0.05 && (double)deltaT < 0.05;                          // :1289
}

// ---------------------------------------------------------------------------------------------------------------
// Per-function parity hooks
// ---------------------------------------------------------------------------------------------------------------
// test hook: cv::solve(DECOMP_QR) of n 6x6 systems by the device routine the solver uses (one thread each)
// (one warp per system: the warp routine of the persistent solver; lane 0 repeats the solve with the one-thread routine of the LM
// hook and raises err_flag = 7 if a single bit differs)
__global__ void __launch_bounds__(64) k_debug_qr6(const float* __restrict__ A, const float* __restrict__ b, int n, float* __restrict__ x, int* err_flag) {
    const int s = blockIdx.x * 2 + (threadIdx.x >> 5);
    if (s >= n) return;
    float xw[6], xs[6];
    qr_solve6_warp(A + 36 * s, b + 6 * s, xw);
    if ((threadIdx.x & 31) == 0) {
        qr_solve6(A + 36 * s, b + 6 * s, xs);
        for (int i = 0; i < 6; ++i) { x[6 * s + i] = xw[i]; if (__float_as_uint(xw[i]) != __float_as_uint(xs[i])) atomicExch(err_flag, 7); }
    }
}

__global__ void __launch_bounds__(S2M_BLOCK) k_surf_optimization(const float4* __restrict__ scan, Count n_scan, const float* __restrict__ tf6,
                                                                const unsigned* __restrict__ cell_start, const float4* __restrict__ gmap,
                                                                GridDims g, Count m_map, float4* __restrict__ coeff_out, unsigned char* __restrict__ flag_out,
                                                                int* __restrict__ idx_out, float* __restrict__ d2_out, float* __restrict__ plane_out,
                                                                float4* __restrict__ sel_out) {
    __shared__ float s_t[12];
    if (threadIdx.x == 0) get_transformation_dev(tf6[3], tf6[4], tf6[5], tf6[0], tf6[1], tf6[2], s_t);
    __syncthreads();
    const int n = n_scan.get();
    const int q = blockIdx.x * S2M_QPB + threadIdx.x / LIORF_G;
    const bool active = q < n;
    float4 ori = active ? scan[q] : make_float4(0, 0, 0, 0);
    float4 sel = apply_affine_dev(s_t, ori);
    Top5 nn; knn5_group(sel, cell_start, gmap, g, nn);
    float4 coeff; float plane[4];
    bool f = surf_point_dev(ori, sel, gmap, nn, coeff, plane) && m_map.get() >= 5;
    if (active && (threadIdx.x & (LIORF_G - 1)) == 0) {
        coeff_out[q] = coeff; flag_out[q] = f ? 1 : 0;
        if (idx_out) for (int j = 0; j < 5; ++j) idx_out[5 * (size_t)q + j] = nn.pos[j] >= 0 ? nn.oi[j] : -1;
        if (d2_out) for (int j = 0; j < 5; ++j) d2_out[5 * (size_t)q + j] = nn.d[j];
        if (plane_out) for (int j = 0; j < 4; ++j) plane_out[4 * (size_t)q + j] = plane[j];
        if (sel_out) sel_out[q] = sel;
    }
}

struct FlagLoad { const unsigned char* f; __device__ __forceinline__ unsigned operator()(int i) const { return f[i] ? 1u : 0u; } };
struct CombineStore {
    const float4* scan; const float4* coeff; float4* ori_out; float4* coeff_out;
    __device__ __forceinline__ void operator()(int i, unsigned v, unsigned excl) const { if (v) { ori_out[excl] = scan[i]; coeff_out[excl] = coeff[i]; } }
};

// LMOptimization hook on the compacted arrays: block-level fp64 reduction into partials, last block solves.
__global__ void __launch_bounds__(256) k_lm_hook(int iter, const float4* __restrict__ ori, const float4* __restrict__ coeff, const unsigned* __restrict__ nsel_dev,
                                                float* tf6, LMDeviceState* st, double* __restrict__ partial, int* counter, float* AtA_out, float* AtB_out,
                                                float* X_out, int* nsel_out, int* conv_out) {
    __shared__ double s_red[8][NPROD];
    __shared__ bool s_last;
    __shared__ float s_A[36], s_V[36];
    const int nsel = (int)*nsel_dev;
    LMTrig trig = lm_trig(tf6);
    double acc[NPROD];
#pragma unroll
    for (int p = 0; p < NPROD; ++p) acc[p] = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nsel; i += gridDim.x * blockDim.x) {
        float v[8]; lm_row_dev(trig, ori[i], coeff[i], v); v[7] = 1.f;
#pragma unroll
        for (int p = 0; p < NPROD; ++p) acc[p] += (double)v[prod_i(p)] * (double)v[prod_j(p)];
    }
#pragma unroll
    for (int p = 0; p < NPROD; ++p) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[p] += __shfl_xor_sync(FULL, acc[p], o);
    }
    if (lane_id() == 0) for (int p = 0; p < NPROD; ++p) s_red[warp_id()][p] = acc[p];
    __syncthreads();
    if (threadIdx.x < NPROD) {
        double s = 0; for (int w = 0; w < 8; ++w) s += s_red[w][threadIdx.x];
        partial[blockIdx.x * NPROD + threadIdx.x] = s;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) { int t = atomicAdd(counter, 1); s_last = (t == (int)gridDim.x - 1); }
    __syncthreads();
    if (!s_last) return;
    __shared__ double s_sum[NPROD];
    if (threadIdx.x < NPROD) {
        const volatile double* vp = partial;
        double s = 0; for (int b = 0; b < (int)gridDim.x; ++b) s += vp[b * NPROD + threadIdx.x];
        s_sum[threadIdx.x] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *counter = 0;
        bool c = lm_solve_dev<false>(iter, s_sum, tf6, st, s_A, s_V, AtA_out, AtB_out, X_out, nsel_out);
        *conv_out = c ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// scan2MapOptimization: ONE persistent kernel (cooperative launch = co-residency guarantee, one CTA per SM) runs the
// whole ≤30-iteration loop with no host round trips.
//
// Per iteration and query (PG = 4 lanes per query):
//   pointSel = T * pointOri
//   candidates: iteration-0 style FULL search walks the 27 cells and, on the way, stores every map point within
//               (1 + m) of the query position q0 (m = S2M_MARGIN) as the query's cached candidate list.
//               While |pointSel - q0| <= m - eps AND pointSel stays in q0's grid cell, every point that can be within
//               1 m of pointSel is in that list (triangle inequality for the radius; same cell so that the unit ball
//               around pointSel stays inside the 27 cells that were searched), so later iterations scan ONLY the
//               ~15-point list — same exact 5-NN, same order.
//   plane     : depends only on the ordered neighbour ids → cached with them; the 5x3 QR is redone only when they change.
//   products  : the 28 normal-equation products are split over the PG lanes, accumulated in fp64.
// Per iteration and CTA: fixed-tree block reduction → partial[cta][28] → arrive on a counter.  The LAST CTA to arrive
// sums the partials in CTA order (deterministic whoever is last), solves the 6x6 system, publishes (pose, converged)
// and releases an epoch flag the other CTAs spin on (bounded).  This replaces grid.sync + 148 redundant sums/solves.
// ---------------------------------------------------------------------------------------------------------------
constexpr int S2MP_BLOCK = 512;             // one CTA per SM
constexpr int S2MP_QPB = 128;               // queries per worker and round with the state in shared memory: one thread each (the first four warps)
constexpr int S2MP_QPB_GLOBAL = S2MP_BLOCK; // ... with the state in global memory: every thread takes a query (no shared-memory capacity to respect)
constexpr int S2MP_WARPS = S2MP_BLOCK / 32;
constexpr int CAND_CAP = 64;                // cached candidates per query (float4 each); overflow → always full search
constexpr float S2M_MARGIN = 0.15f;
// Shared-memory query state of one-round solves (n <= workers x 128: ONE thread serves the same query in every iteration): candidate
// rows of CAND_CAP float4 at an odd stride (65: the eight lanes of a quarter warp, one row each, hit eight different 16-byte bank
// groups), then the four header words of QueryCache as four arrays (lane t reads element t: conflict-free).
constexpr int S2M_ROW = CAND_CAP + 1;
constexpr int S2MP_SMEM = S2MP_QPB * (S2M_ROW + 7) * (int)sizeof(float4);          // 147 456 B (headers: 4 words + a second cached plane of 3)

struct QueryCache {                         // 128 B per query (global-memory layout; shared memory holds the same seven words as seven arrays)
    float4 w[8];                            // 0: q0.xyz, candidate count (-1 none, -2 overflow) | 1: plane A ids 0..3 | 2: id 4, valid, pa, pb |
};                                          // 3: pc, pd, top-5 list positions, plane used last | 4..6: plane B (ids | id 4, valid, pa, pb | pc, pd) | 7: unused
static_assert(sizeof(QueryCache) == 128, "QueryCache must be 128 bytes");
constexpr int S2M_GROW = CAND_CAP + 8;      // global candidate rows: 64 entries + padding for the 8-wide scan

struct alignas(16) S2MMail { float tf[6]; int iters, converged, degenerate, ran, n_scan, n_ds, m_ds, err; int pad[2]; };
static_assert(sizeof(S2MMail) == 64, "S2MMail must be 64 bytes");

constexpr int S2M_MAX_WORKERS = 152;    // >= SMs - 1 (B200: 147)
constexpr int S2M_GT_STRIDE = 160;      // per iteration: [0..W) worker arrival, [156] reducer sums ready, [157] reducer published, [158] worker 0 saw the flag
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// Hand-off words between the workers and the reducer.  Every 64-bit word carries 32 bits of payload and, in its upper half, the epoch
// of the iteration it belongs to (64-bit accesses are single-copy atomic): the data IS the flag.  Neither side needs a release fence
// in front of a separate flag store nor a second round trip behind an acquire load — one L2 round trip per direction and iteration.
__device__ __forceinline__ void st_word2(ulonglong2* p, unsigned long long a, unsigned long long b) {
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ ulonglong2 ld_word2(const ulonglong2* p) {
    ulonglong2 v; asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_word(unsigned long long* p, unsigned long long a) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(a) : "memory"); }
__device__ __forceinline__ unsigned long long ld_word(const unsigned long long* p) {
    unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}

struct S2MArgs {
    const float4* scan; Count n_scan;
    const unsigned* cell_start; const float4* gmap; GridDims g; Count m_map;
    float* tf6;                      // in/out transformTobeMapped (device)
    float tf_init[6]; int use_tf_init;   // the start pose as a kernel argument (saves the one-thread upload kernel in front of the solver)
    LMDeviceState* st;
    ulonglong2* wpart;               // [workers][NPROD]: a worker's partial sums, each fp64 as two epoch-tagged words
    S2MTrace* trace;
    int max_iters; int force_all;
    int no_cache;                    // tests: 1 = never reuse candidate lists / planes (every iteration searches the 27 cells and refits)
    int global_state;                // tests: 1 = per-query state in global memory even when the scan fits one round (the layout of multi-round solves)
    long long* dbg;                  // optional [S2M_MAX_ITERS][8] clock64 phase stamps of CTA 0 (nullptr = off)
    unsigned long long* dbg_gt;      // optional [S2M_MAX_ITERS][S2M_GT_STRIDE] %globaltimer stamps: worker arrivals, flag seen by worker 0, reducer sum-ready / published
    QueryCache* qcache;              // [n_scan bound]   (multi-round solves)
    float4* cand;                    // [n_scan bound][S2M_GROW]
    unsigned long long* res;         // [8]: tf[0..5], converged, nsel — epoch-tagged words written by the reducer
    unsigned epoch_base;             // launch-unique: the epoch of iteration k is epoch_base + k + 1
    int* err_flag;
    S2MMail* mail;                   // optional: everything the host reads after a frame, in ONE 64-byte record (one D2H copy)
    const int* cnt_n_scan;           // device counts copied into the mail (nullable)
};

// ---------------------------------------------------------------------------------------------------------------
// The worker side of an iteration: ONE THREAD PER QUERY, 128 queries per worker and round.
// (Rounds 1 and early 2 gave every query 4 lanes.  ncu showed that pass bound by instruction issue, not memory: 7.6k warp instructions
// per scheduler and iteration, a quarter of them the insertion sort of the per-lane top-5 lists, a tenth their merge, and everything
// after the search executed four times over.)  A round is three phases:
//   1. (thread = query) pointSel; is the cached candidate list still valid (see above)?  If not, queue the query.
//   2. (half warp = queued query) rebuild the list: the 27 cells are walked 64 points at a time and every point within 1 + m of the new
//      q0 is appended (ballot compaction); the half warp leaves the list positions of the five nearest entries.  Steady state: nothing queued.
//   3. (thread = query) exact 5-NN from the list, VERIFY FIRST: the five entries that were nearest last time bound the new 5th-smallest
//      key from above; if exactly those five lie under the bound, still in order, the neighbours are unchanged and nothing is sorted.
//      Otherwise the few entries under the bound are re-ordered (sorting network) or inserted in order.  Then plane reuse (two cached
//      planes per query) / refit, weight, Jacobian row → s_rows.
// followed by the products: 16 chains x 28 products sum their queries' row products in fp64 in a fixed order.
// Same arithmetic per point and the same (distance, index) order as the per-function hooks: neighbour sets and rows are bit-identical.
// Scans that fit ONE round (n <= workers x 128: KITTI, Livox) keep the point in a register and the per-query state in SHARED memory —
// a steady-state iteration touches neither L2 nor HBM; larger scans (OS1-128 at 0.2 m) run several rounds with the same state in
// global memory behind L2 (SM = false), every thread of the CTA taking a query.  A list that overflows CAND_CAP is replaced by the exact
// five nearest points, found by the same half warp (exact_top5_half), and rebuilt in every iteration.
// ---------------------------------------------------------------------------------------------------------------
// (distance, original index) as ONE unsigned 64-bit key: squared distances are non-negative floats, whose bit patterns order like the
// values, and the original indices are non-negative ints — key order == less_di order
__device__ __forceinline__ unsigned long long nn_key(float d, int oi) { return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)oi; }
__device__ __forceinline__ float sqdist_dev(const float4 q, const float4 p) {
    float dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z;
    float d = dx * dx; d += dy * dy; d += dz * dz;                  // FLANN L2_Simple op order, no FMA
    return d;
}

template <bool SM> __device__ __forceinline__ float4 ldq(const float4* p) { if (SM) return *p; return __ldcg(p); }
template <bool SM> __device__ __forceinline__ void stq(float4* p, const float4 v) { if (SM) *p = v; else __stcg(p, v); }

// Sixteen lanes (one half of a warp) per queued query — the two halves of a warp work on two queries at once, so twice as many
// dependent L2 round trips (cell bounds → candidate points) are in flight per SM.  Four candidates per lane and step.
template <bool SM>
__device__ __forceinline__ int build_list_half(const float4 q, const unsigned* __restrict__ cell_start, const float4* __restrict__ gmap, GridDims g,
                                               float4* __restrict__ clist) {
    const int l = lane_id() & 15, hb = lane_id() & 16;
    const unsigned gmask = 0xffffu << hb;
    const int cx = (int)floorf(q.x), cy = (int)floorf(q.y), cz = (int)floorf(q.z);
    const int x0 = (cx - 1) & (g.DX - 1), x1 = cx & (g.DX - 1), x2 = (cx + 1) & (g.DX - 1);
    const float r2c = (1.0f + S2M_MARGIN) * (1.0f + S2M_MARGIN);
    const unsigned lt = (1u << l) - 1u;
    int ncache = 0;
    if (x0 + 2 == x2) {
        // lane r < 9 fetches the bounds of x-run r; an inclusive scan over those lanes turns the lengths into the flattened index space
        unsigned b = 0, len = 0;
        if (l < 9) {
            const int dz = l / 3 - 1, dy = l % 3 - 1;
            const int row = (((cz + dz) & (g.DZ - 1)) * g.DY + ((cy + dy) & (g.DY - 1))) * g.DX;
            b = __ldg(cell_start + row + x0); len = __ldg(cell_start + row + x2 + 1) - b;
        }
        unsigned incl = len;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) { const unsigned v = __shfl_up_sync(gmask, incl, o, 16); if (l >= o) incl += v; }
        unsigned bs[9], P[10];
        P[0] = 0;
#pragma unroll
        for (int r = 0; r < 9; ++r) { bs[r] = __shfl_sync(gmask, b, r, 16); P[r + 1] = __shfl_sync(gmask, incl, r, 16); }
        const unsigned T = P[9];
        for (unsigned f0 = 0; f0 < T; f0 += 64) {
            float4 p[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned f = f0 + 16 * u + l;
                unsigned base = bs[0], pb = 0;
#pragma unroll
                for (int rr = 1; rr < 9; ++rr) if (f >= P[rr]) { base = bs[rr]; pb = P[rr]; }
                p[u] = f < T ? __ldg(gmap + base + (f - pb)) : make_float4(1e30f, 1e30f, 1e30f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float dx = q.x - p[u].x, dy = q.y - p[u].y, dz = q.z - p[u].z;
                float d = dx * dx; d += dy * dy; d += dz * dz;      // FLANN L2_Simple op order, no FMA
                const bool keep = d < r2c;
                const unsigned km = (__ballot_sync(gmask, keep) >> hb) & 0xffffu;
                const int slot = ncache + __popc(km & lt);
                if (keep && slot < CAND_CAP) stq<SM>(clist + slot, p[u]);
                ncache += __popc(km);
            }
        }
    } else {                                     // x-run wraps around the torus: 27 single cells
#pragma unroll 1
        for (int c27 = 0; c27 < 27; ++c27) {
            const int r = c27 / 3, xi = c27 % 3;
            const int dz = r / 3 - 1, dy = r % 3 - 1;
            const int row = (((cz + dz) & (g.DZ - 1)) * g.DY + ((cy + dy) & (g.DY - 1))) * g.DX;
            const int xc = xi == 0 ? x0 : (xi == 1 ? x1 : x2);
            const unsigned b = __ldg(cell_start + row + xc), e = __ldg(cell_start + row + xc + 1);
            for (unsigned f0 = b; f0 < e; f0 += 16) {
                const unsigned f = f0 + l;
                const float4 pt = f < e ? __ldg(gmap + f) : make_float4(1e30f, 1e30f, 1e30f, 0.f);
                float dx = q.x - pt.x, dy2 = q.y - pt.y, dz2 = q.z - pt.z;
                float d = dx * dx; d += dy2 * dy2; d += dz2 * dz2;
                const bool keep = d < r2c;
                const unsigned km = (__ballot_sync(gmask, keep) >> hb) & 0xffffu;
                const int slot = ncache + __popc(km & lt);
                if (keep && slot < CAND_CAP) stq<SM>(clist + slot, pt);
                ncache += __popc(km);
            }
        }
    }
    return ncache <= CAND_CAP ? ncache : -2;
}

// A list that overflows CAND_CAP (dense maps: Livox at 0.3 m leaves has > 64 points within 1.15 m of most queries): the half warp
// searches the 27 cells exactly — per-lane ordered top-5 over its share of the points, merged by (distance, index) — and leaves the
// five winners as the whole list.  Returns how many were found (0..5); such a list is exact for THIS pointSel only.
template <bool SM>
__device__ __forceinline__ int exact_top5_half(const float4 q, const unsigned* __restrict__ cell_start, const float4* __restrict__ gmap, GridDims g,
                                               float4* __restrict__ clist) {
    const int l = lane_id() & 15, hb = lane_id() & 16;
    const unsigned gmask = 0xffffu << hb;
    const int cx = (int)floorf(q.x), cy = (int)floorf(q.y), cz = (int)floorf(q.z);
    Top5 mine; top5_init(mine);
#pragma unroll 1
    for (int c27 = 0; c27 < 27; ++c27) {
        const int dz = c27 / 9 - 1, dy = (c27 / 3) % 3 - 1, dx = c27 % 3 - 1;
        const int cell = (((cz + dz) & (g.DZ - 1)) * g.DY + ((cy + dy) & (g.DY - 1))) * g.DX + ((cx + dx) & (g.DX - 1));
        const unsigned b = __ldg(cell_start + cell), e = __ldg(cell_start + cell + 1);
        for (unsigned k = b + l; k < e; k += 16) {
            const float4 p = __ldg(gmap + k);
            const float d = sqdist_dev(q, p);
            if (d < 1.0f) top5_insert(mine, d, __float_as_int(p.w), (int)k);
        }
    }
    __syncwarp(gmask);
    int found = 0;
#pragma unroll
    for (int r = 0; r < 5; ++r) {                   // 5 rounds of half-warp arg-min by (d, oi); the winning lane pops its head
        float md = mine.d[0]; int mi = mine.oi[0], mp = mine.pos[0];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(gmask, md, o, 16); const int oi = __shfl_xor_sync(gmask, mi, o, 16); const int op = __shfl_xor_sync(gmask, mp, o, 16);
            if (less_di(od, oi, md, mi)) { md = od; mi = oi; mp = op; }
        }
        if (mine.d[0] == md && mine.oi[0] == mi) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { mine.d[j] = mine.d[j + 1]; mine.oi[j] = mine.oi[j + 1]; mine.pos[j] = mine.pos[j + 1]; }
            mine.d[4] = INFINITY; mine.oi[4] = 0x7fffffff; mine.pos[4] = -1;
        }
        if (mp >= 0) { if (l == r) stq<SM>(clist + r, __ldg(gmap + mp)); ++found; }
    }
    return found;
}

struct S2MShared {                      // views into the dynamic shared memory of a one-round worker
    float4* list;                       // [S2MP_QPB][S2M_ROW]
    float4* h;                          // [7][S2MP_QPB]: the seven words of QueryCache as seven arrays (lane t reads element t: conflict-free)
};

// phase 2 epilogue, one half warp: list positions of the five smallest keys of a freshly built list (cnt >= 5), packed 6 bits each,
// smallest first — the prior that phase 3 verifies from now on.  Lane l holds entries l, l + 16, l + 32, l + 48.
template <bool SM>
__device__ __forceinline__ unsigned top5_positions_half(const float4 q, const float4* __restrict__ row, int cnt) {
    const int l = lane_id() & 15, hb = lane_id() & 16;
    const unsigned gmask = 0xffffu << hb;
    unsigned long long k4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        k4[u] = ~0ull;
        if (l + 16 * u < cnt) { const float4 p = ldq<SM>(row + l + 16 * u); k4[u] = nn_key(sqdist_dev(q, p), __float_as_int(p.w)); }
    }
    unsigned packed = 0;
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        unsigned long long mine = k4[0]; int mu = 0;
#pragma unroll
        for (int u = 1; u < 4; ++u) if (k4[u] < mine) { mine = k4[u]; mu = u; }
        unsigned long long m = mine;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) { const unsigned long long v = __shfl_xor_sync(gmask, m, o, 16); m = v < m ? v : m; }
        const int owner = __ffs((__ballot_sync(gmask, mine == m) >> hb) & 0xffffu) - 1;          // keys are unique (distinct original indices)
        packed |= (unsigned)__shfl_sync(gmask, l + 16 * mu, owner, 16) << (6 * r);
        if (l == owner) {
#pragma unroll
            for (int u = 0; u < 4; ++u) if (u == mu) k4[u] = ~0ull;
        }
    }
    return packed;
}

template <bool SM>
__device__ __forceinline__ void s2m_round(const S2MArgs& a, const int n, const int W, const int iter, const int round, const float (&s_t)[12], const LMTrig& s_trig,
                                          float (*s_rows)[8], const S2MShared& S, const float4 ori_keep, const float rr_keep, int* s_todo, int* s_ntodo, long long* dbg_p2) {
    const int tid = threadIdx.x;
    // slot s of this round is query (round * 128 + s) * W + worker: dense and sparse regions of the scan spread over all SMs
    constexpr int QPB = SM ? S2MP_QPB : S2MP_QPB_GLOBAL;
    auto query_of = [&](int slot) { return (round * QPB + slot) * W + (int)blockIdx.x; };
    auto hdr = [&](int slot, int k) -> float4* { if (SM) return S.h + k * S2MP_QPB + slot; return a.qcache[query_of(slot)].w + k; };
    auto rowp = [&](int slot) -> float4* { if (SM) return S.list + slot * S2M_ROW; return a.cand + (size_t)query_of(slot) * S2M_GROW; };
    const bool mine = tid < QPB && query_of(tid) < n;
    float4 ori = ori_keep; float rr = rr_keep;
    if (!SM && mine) { ori = __ldg(a.scan + query_of(tid)); rr = sqrtf(sqrtf(ori.x * ori.x + ori.y * ori.y + ori.z * ori.z)); }
    float4 sel = make_float4(0.f, 0.f, 0.f, 0.f);
    // ---- 1. pointSel; queue the queries whose candidate list must be rebuilt ----
    if (mine) {
        sel = apply_affine_dev(s_t, ori);
        const float4 h0 = ldq<SM>(hdr(tid, 0));
        const int cnt = __float_as_int(h0.w);
        float mx = sel.x - h0.x, my = sel.y - h0.y, mz = sel.z - h0.z;
        float moved = mx * mx; moved += my * my; moved += mz * mz;
        const float lim = (S2M_MARGIN - 1e-3f) * (S2M_MARGIN - 1e-3f);
        // the list holds the points within 1 + m of q0 THAT LIE IN q0's 27 cells; the unit ball around pointSel stays
        // inside that block of cells only while pointSel is in q0's own cell
        const bool same_cell = floorf(sel.x) == floorf(h0.x) && floorf(sel.y) == floorf(h0.y) && floorf(sel.z) == floorf(h0.z);
        const bool valid = iter > 0 && cnt >= 0 && moved <= lim && same_cell && !a.no_cache;
        if (!valid) { s_todo[atomicAdd(s_ntodo, 1)] = tid; stq<SM>(hdr(tid, 0), make_float4(sel.x, sel.y, sel.z, __int_as_float(-1))); }
        if (a.dbg_gt && !valid && iter >= 10) atomicAdd(a.dbg_gt + 40 * S2M_GT_STRIDE + blockIdx.x, 1ull);
        if (a.dbg_gt && !valid && iter >= 10 && h0.x != h0.x) atomicAdd(a.dbg_gt + 43 * S2M_GT_STRIDE + blockIdx.x, 1ull);     // overflowed list
        // iteration 0 never trusts a cached plane (it belongs to an earlier launch / another map)
        if (iter == 0) {
            const float4 none = make_float4(__int_as_float(-1), __int_as_float(-1), __int_as_float(-1), __int_as_float(-1));
            stq<SM>(hdr(tid, 1), none); stq<SM>(hdr(tid, 4), none);
        }
    }
    __syncthreads();
    // ---- 2. one half warp per queued query rebuilds its list around the new q0 and leaves the positions of its five nearest entries ----
    const int ntodo = *s_ntodo;
    for (int k = 2 * warp_id() + (lane_id() >> 4); k < ntodo; k += 2 * S2MP_WARPS) {       // the two halves of a warp take one query each
        const int slot = s_todo[k];
        const float4 q0 = ldq<SM>(hdr(slot, 0));
        float4* row = rowp(slot);
        int c = build_list_half<SM>(q0, a.cell_start, a.gmap, a.g, row);
        __syncwarp(0xffffu << (lane_id() & 16));
        unsigned packed = 0;
        float q0x = q0.x;
        if (c == -2) {                               // overflow: the five nearest ARE the list, in order; NaN in q0.x keeps it from ever being reused
            c = exact_top5_half<SM>(q0, a.cell_start, a.gmap, a.g, row);
            __syncwarp(0xffffu << (lane_id() & 16));
            packed = 0u | (1u << 6) | (2u << 12) | (3u << 18) | (4u << 24);
            q0x = __int_as_float(0x7fc00000);
        } else if (c >= 5) packed = top5_positions_half<SM>(q0, row, c);
        if ((lane_id() & 15) == 0) {
            stq<SM>(hdr(slot, 0), make_float4(q0x, q0.y, q0.z, __int_as_float(c)));
            float4 h3 = ldq<SM>(hdr(slot, 3)); h3.z = __uint_as_float(packed); stq<SM>(hdr(slot, 3), h3);
        }
    }
    __syncthreads();
    if (dbg_p2) *dbg_p2 = clock64();
    if (tid == 0) *s_ntodo = 0;                  // the next phase 1 is at least two barriers away
    // ---- 3. exact 5-NN from the list, plane, weight, Jacobian row ----
    if (tid < QPB) {
        bool f = false; float4 coeff = make_float4(0.f, 0.f, 0.f, 0.f);
        float v[8];
        if (mine) {
            const int cnt = __float_as_int(ldq<SM>(hdr(tid, 0)).w);
            const float4* row = rowp(tid);
            const float4 h3 = ldq<SM>(hdr(tid, 3));
            Top5 nn; top5_init(nn);
            if (cnt >= 5) {
                // The five entries that were nearest last time (or when the list was built) bound the new 5th-smallest key from above:
                // only entries with key <= thr can be among the new five.  Usually those are the same five in the same order — then the
                // neighbours are unchanged and nothing is sorted at all; otherwise the few entries under the bound are re-ordered / inserted.
                unsigned pk = __float_as_uint(h3.z);
                unsigned long long ck[5]; float cd[5]; int coi[5], cpos[5];
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    cpos[j] = (int)((pk >> (6 * j)) & 63u);
                    const float4 p = ldq<SM>(row + cpos[j]);
                    cd[j] = sqdist_dev(sel, p); coi[j] = __float_as_int(p.w); ck[j] = nn_key(cd[j], coi[j]);
                }
                unsigned long long thr = ck[0];
#pragma unroll
                for (int j = 1; j < 5; ++j) thr = ck[j] > thr ? ck[j] : thr;
                unsigned long long mask = 0; int n1 = 0;                 // n1: entries within 1 m — fewer than five ⇒ no plane (:1097), nothing to order
                for (int f0 = 0; f0 < cnt; f0 += 8) {
                    float4 p[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) p[u] = ldq<SM>(row + f0 + u);          // the row is padded: entries past cnt are read but masked out
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const float d = sqdist_dev(sel, p[u]);
                        const unsigned long long k = nn_key(d, __float_as_int(p[u].w));
                        mask |= (unsigned long long)((k <= thr) && (f0 + u < cnt)) << (f0 + u);
                        n1 += (d < 1.0f) && (f0 + u < cnt);
                    }
                }
                const bool sorted = ck[0] < ck[1] && ck[1] < ck[2] && ck[2] < ck[3] && ck[3] < ck[4];
                if (n1 < 5) {
                    // nn stays empty
                } else if (__popcll(mask) == 5 && sorted && cd[4] < 1.0f) {
#pragma unroll
                    for (int j = 0; j < 5; ++j) { nn.d[j] = cd[j]; nn.oi[j] = coi[j]; nn.pos[j] = cpos[j]; }
                } else if (__popcll(mask) <= 6) {
                    // the same five in a different order, or ONE entry slipped under the bound (n1 >= 5, so the five smallest keys are all
                    // within 1 m): a 9-exchange sorting network on the cached keys, then the newcomer takes its place and the last one drops out
#define S2M_CE(i, j) if (ck[j] < ck[i]) { const unsigned long long tk = ck[i]; ck[i] = ck[j]; ck[j] = tk; const float td = cd[i]; cd[i] = cd[j]; cd[j] = td; \
                                           const int to = coi[i]; coi[i] = coi[j]; coi[j] = to; const int tp = cpos[i]; cpos[i] = cpos[j]; cpos[j] = tp; }
                    S2M_CE(0, 1) S2M_CE(3, 4) S2M_CE(2, 4) S2M_CE(2, 3) S2M_CE(1, 4) S2M_CE(0, 3) S2M_CE(0, 2) S2M_CE(1, 3) S2M_CE(1, 2)
#undef S2M_CE
                    unsigned long long others = mask;
#pragma unroll
                    for (int j = 0; j < 5; ++j) others &= ~(1ull << cpos[j]);
                    if (others) {
                        const int ip = __ffsll((long long)others) - 1;
                        const float4 p = ldq<SM>(row + ip);
                        float nd = sqdist_dev(sel, p); int noi = __float_as_int(p.w), np = ip;
                        unsigned long long nk = nn_key(nd, noi);
#pragma unroll
                        for (int j = 0; j < 5; ++j) {                 // ripple insert: the larger of (newcomer, entry j) moves on
                            if (nk < ck[j]) {
                                const unsigned long long tk = ck[j]; ck[j] = nk; nk = tk; const float td = cd[j]; cd[j] = nd; nd = td;
                                const int to = coi[j]; coi[j] = noi; noi = to; const int tp = cpos[j]; cpos[j] = np; np = tp;
                            }
                        }
                    }
                    pk = 0;
#pragma unroll
                    for (int j = 0; j < 5; ++j) { nn.d[j] = cd[j]; nn.oi[j] = coi[j]; nn.pos[j] = cpos[j]; pk |= (unsigned)cpos[j] << (6 * j); }
                    float4 h3n = h3; h3n.z = __uint_as_float(pk); stq<SM>(hdr(tid, 3), h3n);
                } else {
                    if (a.dbg_gt && iter >= 10) atomicAdd(a.dbg_gt + 41 * S2M_GT_STRIDE + blockIdx.x, 1ull);
                    // ordered insertion of the entries under the bound, on the u64 keys (distance and index travel inside the key)
                    unsigned long long tk[5] = {~0ull, ~0ull, ~0ull, ~0ull, ~0ull}; int tp[5] = {-1, -1, -1, -1, -1};
                    while (mask) {
                        const int i = __ffsll((long long)mask) - 1; mask &= mask - 1;
                        const float4 p = ldq<SM>(row + i);
                        const float d = sqdist_dev(sel, p);
                        unsigned long long k = nn_key(d, __float_as_int(p.w)); int kp = i;
                        if (d < 1.0f && k < tk[4]) {
#pragma unroll
                            for (int j = 0; j < 5; ++j) if (k < tk[j]) { const unsigned long long t = tk[j]; tk[j] = k; k = t; const int t2 = tp[j]; tp[j] = kp; kp = t2; }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 5; ++j) if (tp[j] >= 0) { nn.d[j] = __uint_as_float((unsigned)(tk[j] >> 32)); nn.oi[j] = (int)(unsigned)tk[j]; nn.pos[j] = tp[j]; }
                    if (nn.pos[4] != -1) {
                        pk = 0;
#pragma unroll
                        for (int j = 0; j < 5; ++j) pk |= (unsigned)nn.pos[j] << (6 * j);
                        float4 h3n = h3; h3n.z = __uint_as_float(pk); stq<SM>(hdr(tid, 3), h3n);
                    }
                }
            }
            const bool have5 = nn.pos[4] != -1 && (double)nn.d[4] < 1.0;                // :1097
            if (have5) {
                // ---- plane: reuse when the ordered neighbour ids match one of the two cached planes ----
                const float4 h1 = ldq<SM>(hdr(tid, 1)), h2 = ldq<SM>(hdr(tid, 2)), g1 = ldq<SM>(hdr(tid, 4)), g2 = ldq<SM>(hdr(tid, 5));
                const bool use = iter > 0 && !a.no_cache;
                const bool same0 = use && __float_as_int(h1.x) == nn.oi[0] && __float_as_int(h1.y) == nn.oi[1] && __float_as_int(h1.z) == nn.oi[2] &&
                                   __float_as_int(h1.w) == nn.oi[3] && __float_as_int(h2.x) == nn.oi[4];
                const bool same1 = use && __float_as_int(g1.x) == nn.oi[0] && __float_as_int(g1.y) == nn.oi[1] && __float_as_int(g1.z) == nn.oi[2] &&
                                   __float_as_int(g1.w) == nn.oi[3] && __float_as_int(g2.x) == nn.oi[4];
                float pa, pb, pc, pd; bool planeValid;
                float4 h3c = ldq<SM>(hdr(tid, 3));                                       // (the top-5 positions may just have been rewritten)
                if (same0) { planeValid = __float_as_int(h2.y) != 0; pa = h2.z; pb = h2.w; pc = h3c.x; pd = h3c.y; }
                else if (same1) { const float4 g3 = ldq<SM>(hdr(tid, 6)); planeValid = __float_as_int(g2.y) != 0; pa = g2.z; pb = g2.w; pc = g3.x; pd = g3.y; }
                else {
                    float A[5][3];
#pragma unroll
                    for (int j = 0; j < 5; ++j) {
                        const float4 mpt = ldq<SM>(row + nn.pos[j]);
                        A[j][0] = mpt.x; A[j][1] = mpt.y; A[j][2] = mpt.z;
                    }
                    float x[3];
                    if (a.dbg_gt && iter >= 10) atomicAdd(a.dbg_gt + 42 * S2M_GT_STRIDE + blockIdx.x, 1ull);
                    colpiv_qr_solve_5x3(A, x);                                           // :1104
                    pa = x[0]; pb = x[1]; pc = x[2]; pd = 1.f;
                    float ps = sqrtf(pa * pa + pb * pb + pc * pc);                       // :1111
                    pa /= ps; pb /= ps; pc /= ps; pd /= ps;
                    planeValid = true;
#pragma unroll
                    for (int j = 0; j < 5; ++j)                                          // :1115-1122
                        if ((double)fabsf(pa * A[j][0] + pb * A[j][1] + pc * A[j][2] + pd) > 0.2) planeValid = false;
                    // replace the plane that was NOT used last
                    const bool into1 = __float_as_int(h3c.w) == 0;
                    const float4 n1v = make_float4(__int_as_float(nn.oi[0]), __int_as_float(nn.oi[1]), __int_as_float(nn.oi[2]), __int_as_float(nn.oi[3]));
                    const float4 n2v = make_float4(__int_as_float(nn.oi[4]), __int_as_float(planeValid ? 1 : 0), pa, pb);
                    if (into1) { stq<SM>(hdr(tid, 4), n1v); stq<SM>(hdr(tid, 5), n2v); stq<SM>(hdr(tid, 6), make_float4(pc, pd, 0.f, 0.f)); h3c.w = __int_as_float(1); }
                    else { stq<SM>(hdr(tid, 1), n1v); stq<SM>(hdr(tid, 2), n2v); h3c.x = pc; h3c.y = pd; h3c.w = __int_as_float(0); }
                    stq<SM>(hdr(tid, 3), h3c);
                }
                if (same0 || same1) {
                    const int used = same0 ? 0 : 1;
                    if (__float_as_int(h3c.w) != used) { h3c.w = __int_as_float(used); stq<SM>(hdr(tid, 3), h3c); }
                }
                if (planeValid) {
                    float pd2 = pa * sel.x + pb * sel.y + pc * sel.z + pd;               // :1125
                    float sw = (float)(1.0 - 0.9 * (double)fabsf(pd2) / (double)rr);     // :1127-1128 (rr = sqrt(sqrt(|pointOri|^2)) does not change between iterations)
                    coeff = make_float4(sw * pa, sw * pb, sw * pc, sw * pd2);            // :1130-1133
                    f = (double)sw > 0.1;                                                // :1135
                }
            }
            lm_row_dev(s_trig, ori, coeff, v); v[7] = 1.f;
        }
        float4* r4 = reinterpret_cast<float4*>(&s_rows[tid][0]);             // zero row when the point is not selected / the slot is idle
        r4[0] = f ? make_float4(v[0], v[1], v[2], v[3]) : make_float4(0.f, 0.f, 0.f, 0.f);
        r4[1] = f ? make_float4(v[4], v[5], v[6], v[7]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// The reducer's solve (lm_solve_dev's arithmetic) by ONE WARP: the Householder sweep over the six columns and the right-hand side is
// the long pole of an iteration, and its columns are independent (qr_solve6_warp).  sums: 28 fp64 totals in shared memory.
// Every lane returns the same (tf, converged, nsel).
template <bool FAST_DEGENERACY>
__device__ __forceinline__ bool lm_solve_warp(int iter, const double* sums, float (&tf)[6], LMDeviceState* st, float* scratchA /*36*/, float* scratchV /*36*/, int& nsel) {
    float X[6];
    nsel = (int)(sums[27] + 0.5);
    if (nsel < 50) return false;                                                    // :1178
    {   // lane j < 6 takes column j of AtA (symmetric: entry (i, j) is product rs(min) + |i - j|), lane 6 takes AtB
        const int j = lane_id();
        float col[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const int lo = i < j ? i : j, hi = i < j ? j : i;
            const int p = j < 6 ? lo * 6 - lo * (lo - 1) / 2 + (hi - lo) : (j == 6 ? 21 + i : 27);
            col[i] = (float)sums[p];
        }
        qr_solve6_warp_cols(col, X);                                                // :1240
    }
    if (iter == 0) {                                                                // :1242-1264
        float AtA[36];
#pragma unroll
        for (int p = 0; p < 21; ++p) { const float v = (float)sums[p]; AtA[prod_i(p) * 6 + prod_j(p)] = v; AtA[prod_j(p) * 6 + prod_i(p)] = v; }
        const bool cert = FAST_DEGENERACY && certify_non_degenerate(AtA);           // same verdict in every lane
        if (lane_id() == 0) {
            if (cert) {
                st->isDegenerate = 0;
#pragma unroll
                for (int i = 0; i < 36; ++i) st->matP[i] = (i % 7 == 0) ? 1.f : 0.f;
            } else {
                float tmpA[36];
#pragma unroll
                for (int i = 0; i < 36; ++i) tmpA[i] = AtA[i];
                lm_degeneracy_dev(tmpA, st, scratchA, scratchV);
            }
        }
        __syncwarp();
    }
    if (st->isDegenerate) {                                                         // :1266-1271
        float X2[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) X2[i] = X[i];
        gemv6(st->matP, X2, X);
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) tf[i] += X[i];                                      // :1273-1278
    const float r2d = 57.29578f;
    double a0 = (double)(X[0] * r2d), a1 = (double)(X[1] * r2d), a2 = (double)(X[2] * r2d);
    double t0 = (double)(X[3] * 100), t1 = (double)(X[4] * 100), t2 = (double)(X[5] * 100);
    float deltaR = (float)sqrt(a0 * a0 + a1 * a1 + a2 * a2);                        // :1280-1287
    float deltaT = (float)sqrt(t0 * t0 + t1 * t1 + t2 * t2);
    return (double)deltaR < 0.05 && (double)deltaT < 0.05;                          // :1289
}

// ---------------------------------------------------------------------------------------------------------------
// scan2MapOptimization: ONE persistent kernel (cooperative launch = co-residency guarantee, one CTA per SM) runs the
// whole ≤30-iteration loop with no host round trips.
//
// Per iteration and query (PG = 4, 8 or 16 lanes per query):
//   pointSel = T * pointOri
//   candidates: iteration-0 style FULL search walks the 27 cells and, on the way, stores every map point within
//               (1 + m) of the query position q0 (m = S2M_MARGIN) as the query's cached candidate list.
//               While |pointSel - q0| <= m - eps AND pointSel stays in q0's grid cell, every point that can be within
//               1 m of pointSel is in that list (triangle inequality for the radius; same cell so that the unit ball
//               around pointSel stays inside the 27 cells that were searched), so later iterations scan ONLY the
//               ~15-point list — same exact 5-NN, same order.
//   plane     : depends only on the ordered neighbour ids → cached with them; the 5x3 QR is redone only when they change.
//   products  : the 28 normal-equation products are split over four lanes of the group, accumulated in fp64.
// Per iteration and CTA: fixed-tree block reduction → 28 epoch-tagged partial sums written straight to the reducer's inbox.
// The REDUCER CTA collects the inboxes as they fill, sums them in worker order (deterministic), one of its warps solves the 6x6
// system and publishes (pose, converged, nsel) as eight epoch-tagged words the workers poll.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(S2MP_BLOCK, 1) k_scan2map_persistent(S2MArgs a) {
    extern __shared__ __align__(16) float4 s_state[];     // S2MP_SMEM bytes: per-query candidate rows + headers (one-round solves)
    __shared__ int s_todo[S2MP_QPB_GLOBAL];
    __shared__ int s_ntodo;
    __shared__ float s_tf[6];
    __shared__ float s_sc[6];                              // cos/sin of yaw, pitch, roll for this iteration
    __shared__ __align__(16) float s_rows[S2MP_QPB_GLOBAL][8];
    __shared__ double s_red[S2MP_WARPS][NPROD];
    __shared__ double s_sum[NPROD];
    __shared__ float s_A[36], s_V[36];
    __shared__ LMDeviceState s_st;
    __shared__ int s_conv;

    // Roles: CTAs 0 .. W-1 are WORKERS (queries), the last CTA is the REDUCER: it owns the LM state, sums the workers'
    // partials in CTA order as they arrive, solves the 6x6 system and publishes (pose, converged).  A fixed reducer keeps the
    // solve's code and state hot in ONE SM's caches — with "the last CTA to arrive solves" a different SM ran ~60 KB of cold
    // straight-line code every iteration (measured: 2.4 us warm, 5.6 us cold) — and its sum is done by the time the slowest worker
    // arrives instead of starting there.
    const int W = (int)gridDim.x - 1;
    const bool reducer = (int)blockIdx.x == W;
    const int n = a.n_scan.get();
    const int m = a.m_map.get();
    const bool one_round = !a.global_state && n <= W * S2MP_QPB;              // query state in registers / shared memory
    const int qpb = one_round ? S2MP_QPB : S2MP_QPB_GLOBAL;
    const int rounds = one_round ? 1 : (n + W * qpb - 1) / (W * qpb);
    S2MShared S; S.list = s_state; S.h = s_state + S2MP_QPB * S2M_ROW;
    if (threadIdx.x < 6) s_tf[threadIdx.x] = a.use_tf_init ? a.tf_init[threadIdx.x] : a.tf6[threadIdx.x];
    if (threadIdx.x == 32) { s_conv = 0; s_ntodo = 0; }
    if (reducer && threadIdx.x >= 64 && threadIdx.x < 64 + 37)          // persistent LM state (members :139-140)
        reinterpret_cast<int*>(&s_st)[threadIdx.x - 64] = __ldcg(reinterpret_cast<const int*>(a.st) + (threadIdx.x - 64));
    float4 ori_keep = make_float4(0.f, 0.f, 0.f, 0.f);
    float rr_keep = 0.f;
    if (!reducer && one_round && threadIdx.x < S2MP_QPB) { const int q = (int)threadIdx.x * W + (int)blockIdx.x; if (q < n) { ori_keep = __ldg(a.scan + q); rr_keep = sqrtf(sqrtf(ori_keep.x * ori_keep.x + ori_keep.y * ori_keep.y + ori_keep.z * ori_keep.z)); } }
    // product owned by this thread in the one-round reduction: chain k16 = tid / 28 of 16, product p16 = tid % 28
    const int k16 = (int)threadIdx.x / NPROD, p16 = (int)threadIdx.x % NPROD;
    const int pi16 = prod_i(p16), pj16 = prod_j(p16);
    __syncthreads();
    // guards of scan2MapOptimization (:1297-1300): a map must exist (and hold >= 5 points for the 5-NN), n > 30
    const bool run = (m >= 5) && (n > 30);
    int iters_done = 0, converged = 0;
    if (run) {
        // iteration 0 never trusts the per-query caches (they belong to an earlier launch / another map)
        for (int iter = 0; iter < a.max_iters; ++iter) {
            const unsigned epoch = a.epoch_base + (unsigned)iter + 1u;
            const unsigned long long etag = (unsigned long long)epoch << 32;
            if (!reducer) {
                const bool dbg = a.dbg && blockIdx.x == 0 && threadIdx.x == 0 && iter < S2M_MAX_ITERS;
                if (dbg) a.dbg[iter * 8 + 0] = clock64();
                // six fp64-rounded trig values in parallel: (cos, sin) of yaw = tf[2], pitch = tf[1], roll = tf[0]
                if (threadIdx.x < 6) { const float ang = s_tf[2 - (threadIdx.x >> 1)]; s_sc[threadIdx.x] = (threadIdx.x & 1) ? sin_f(ang) : cos_f(ang); }
                __syncthreads();
                float s_t[12]; LMTrig s_trig;
                {   // pcl::getTransformation from the shared trig values (same arithmetic as get_transformation_dev)
                    const float A = s_sc[0], B = s_sc[1], Cc = s_sc[2], D = s_sc[3], E = s_sc[4], F = s_sc[5], DE = D * E, DF = D * F;
                    s_t[0] = A * Cc; s_t[1] = A * DF - B * E; s_t[2] = B * F + A * DE; s_t[3] = s_tf[3];
                    s_t[4] = B * Cc; s_t[5] = A * E + B * DF; s_t[6] = B * DE - A * F; s_t[7] = s_tf[4];
                    s_t[8] = -D;     s_t[9] = Cc * F;         s_t[10] = Cc * E;        s_t[11] = s_tf[5];
                    s_trig.srx = B; s_trig.crx = A; s_trig.sry = D; s_trig.cry = Cc; s_trig.srz = F; s_trig.crz = E;   // :1170-1175
                }
                double sacc = 0.0;
                for (int round = 0; round < rounds; ++round) {
                    long long* dbg_p2 = dbg && round == 0 ? a.dbg + iter * 8 + 3 : nullptr;
                    if (one_round) s2m_round<true>(a, n, W, iter, round, s_t, s_trig, s_rows, S, ori_keep, rr_keep, s_todo, &s_ntodo, dbg_p2);
                    else s2m_round<false>(a, n, W, iter, round, s_t, s_trig, s_rows, S, ori_keep, rr_keep, s_todo, &s_ntodo, dbg_p2);
                    if (dbg && round == rounds - 1) a.dbg[iter * 8 + 1] = clock64();
                    __syncthreads();
                    if (k16 < S2MP_WARPS) {               // 16 chains x 28 products, chain k takes this round's queries k, k + 16, ... in ascending order
                        const int nq = max(0, min(qpb, (n - (int)blockIdx.x + W - 1) / W - round * qpb));
                        for (int q = k16; q < nq; q += S2MP_WARPS) sacc += (double)s_rows[q][pi16] * (double)s_rows[q][pj16];
                    }
                    if (round + 1 < rounds) __syncthreads();      // the next round overwrites the rows
                }
                if (k16 < S2MP_WARPS) s_red[k16][p16] = sacc;
                __syncthreads();
                if (threadIdx.x < NPROD) {
                    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
                    for (int w = 0; w < S2MP_WARPS; w += 4) { s0 += s_red[w][threadIdx.x]; s1 += s_red[w + 1][threadIdx.x]; s2 += s_red[w + 2][threadIdx.x]; s3 += s_red[w + 3][threadIdx.x]; }
                    const unsigned long long bits = (unsigned long long)__double_as_longlong((s0 + s1) + (s2 + s3));
                    st_word2(a.wpart + (size_t)blockIdx.x * NPROD + threadIdx.x, (bits & 0xffffffffull) | etag, (bits >> 32) | etag);
                }
                if (dbg) a.dbg[iter * 8 + 2] = clock64();
                if (threadIdx.x == 0 && a.dbg_gt && iter < S2M_MAX_ITERS) a.dbg_gt[iter * S2M_GT_STRIDE + blockIdx.x] = gtimer();
                // the reducer's answer: eight lanes poll one word each (one L2 round trip once it is there)
                if (threadIdx.x < 8) {
                    unsigned long long v; unsigned spins = 0;
                    while (true) {
                        v = ld_word(a.res + threadIdx.x);
                        if ((unsigned)(v >> 32) == epoch) break;
                        if (++spins > (1u << 26)) { atomicExch(a.err_flag, 2); break; }
                    }
                    if (threadIdx.x < 6) s_tf[threadIdx.x] = __uint_as_float((unsigned)v);
                    else if (threadIdx.x == 6) s_conv = (int)(unsigned)v;
                    if (dbg) a.dbg[iter * 8 + 4] = clock64();
                    if (a.dbg_gt && blockIdx.x == 0 && threadIdx.x == 0 && iter < S2M_MAX_ITERS) a.dbg_gt[iter * S2M_GT_STRIDE + 158] = gtimer();
                }
                __syncthreads();
                if (dbg) a.dbg[iter * 8 + 5] = clock64();
            } else {
                // ---- reducer: warp k polls the inboxes of workers k, k + 16, ... (lane = product; an entry is complete when both halves
                // carry this iteration's epoch; only the missing ones are re-read) and sums them in that FIXED order; the 16 chunk sums are
                // combined by the 4-chain tree below, so the result does not depend on the arrival order ----
                {
                    constexpr int MAXB = (S2M_MAX_WORKERS + S2MP_WARPS - 1) / S2MP_WARPS;
                    const int wk = warp_id(), pl = lane_id();
                    ulonglong2 w2[MAXB]; bool got[MAXB];
#pragma unroll
                    for (int j = 0; j < MAXB; ++j) got[j] = !(wk + S2MP_WARPS * j < W && pl < NPROD);
                    unsigned spins = 0;
                    while (true) {
#pragma unroll
                        for (int j = 0; j < MAXB; ++j) if (!got[j]) w2[j] = ld_word2(a.wpart + (size_t)(wk + S2MP_WARPS * j) * NPROD + pl);
                        bool all = true;
#pragma unroll
                        for (int j = 0; j < MAXB; ++j) {
                            if (got[j]) continue;
                            if ((unsigned)(w2[j].x >> 32) == epoch && (unsigned)(w2[j].y >> 32) == epoch) got[j] = true; else all = false;
                        }
                        if (all) break;
                        if (++spins > (1u << 24)) { atomicExch(a.err_flag, 2); break; }
                    }
                    if (pl < NPROD) {
                        double s0 = 0;
#pragma unroll
                        for (int j = 0; j < MAXB; ++j)
                            if (wk + S2MP_WARPS * j < W) s0 += __longlong_as_double((long long)((w2[j].x & 0xffffffffull) | (w2[j].y << 32)));
                        s_red[wk][pl] = s0;
                    }
                }
                __syncthreads();
                const long long lc0 = clock64();
                if (threadIdx.x < NPROD) {
                    double t0 = 0, t1 = 0, t2 = 0, t3 = 0;
#pragma unroll
                    for (int ww = 0; ww < S2MP_WARPS; ww += 4) { t0 += s_red[ww][threadIdx.x]; t1 += s_red[ww + 1][threadIdx.x]; t2 += s_red[ww + 2][threadIdx.x]; t3 += s_red[ww + 3][threadIdx.x]; }
                    s_sum[threadIdx.x] = (t0 + t1) + (t2 + t3);
                }
                __syncthreads();
                const long long lc1 = clock64();
                if (threadIdx.x < 32) {
                    if (threadIdx.x == 0 && a.dbg_gt && iter < S2M_MAX_ITERS) a.dbg_gt[iter * S2M_GT_STRIDE + 156] = gtimer();
                    int nsel = 0;
                    float tfn[6];
#pragma unroll
                    for (int k = 0; k < 6; ++k) tfn[k] = s_tf[k];
                    const bool c = lm_solve_warp<true>(iter, s_sum, tfn, &s_st, s_A, s_V, nsel);
                    if (threadIdx.x < 8) {          // the answer: eight self-validating words
                        unsigned pay = 0;
#pragma unroll
                        for (int k = 0; k < 6; ++k) if ((int)threadIdx.x == k) pay = __float_as_uint(tfn[k]);
                        if (threadIdx.x == 6) pay = c ? 1u : 0u;
                        if (threadIdx.x == 7) pay = (unsigned)nsel;
                        st_word(a.res + threadIdx.x, (unsigned long long)pay | etag);
                    }
                    if (threadIdx.x == 0) {
                        if (a.dbg_gt && iter < S2M_MAX_ITERS) a.dbg_gt[iter * S2M_GT_STRIDE + 157] = gtimer();
                        // off the critical path: the workers are already running the next iteration
#pragma unroll
                        for (int k = 0; k < 6; ++k) s_tf[k] = tfn[k];
                        s_conv = c ? 1 : 0;
                        if (iter == 0) *a.st = s_st;
                        if (a.dbg && iter < S2M_MAX_ITERS) { a.dbg[iter * 8 + 6] = lc1 - lc0; a.dbg[iter * 8 + 7] = clock64() - lc1; }
                        if (a.trace && iter < S2M_MAX_ITERS) {
                            for (int k = 0; k < 6; ++k) a.trace->pose[iter][k] = tfn[k];
                            a.trace->nsel[iter] = nsel;
                            a.trace->degenerate = s_st.isDegenerate;
                        }
                    }
                }
                __syncthreads();
            }
            iters_done = iter + 1;
            if (s_conv) { converged = 1; if (!a.force_all) break; }
        }
    }
    if (reducer && threadIdx.x == 0) {
        for (int k = 0; k < 6; ++k) a.tf6[k] = s_tf[k];
        if (a.trace) { a.trace->iters = iters_done; a.trace->converged = converged; a.trace->ran = run ? 1 : 0; if (!run) a.trace->degenerate = s_st.isDegenerate; }
        if (a.mail) {
            S2MMail mm;
            for (int k = 0; k < 6; ++k) mm.tf[k] = s_tf[k];
            mm.iters = iters_done; mm.converged = converged; mm.degenerate = s_st.isDegenerate; mm.ran = run ? 1 : 0;
            mm.n_scan = a.cnt_n_scan ? __ldcg(a.cnt_n_scan) : -1; mm.n_ds = n; mm.m_ds = m; mm.err = __ldcg(a.err_flag); mm.pad[0] = mm.pad[1] = 0;
            *a.mail = mm;
        }
    }
}

}  // namespace liorf

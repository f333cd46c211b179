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
constexpr int S2M_BLOCK = 256;
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
__device__ __forceinline__ void knn5_group(const float4 q, const unsigned* __restrict__ cell_start, const float4* __restrict__ gmap,
                                           GridDims g, Top5& res) {
    const int gl = threadIdx.x & (LIORF_G - 1);
    Top5 mine; top5_init(mine);
    const int cx = (int)floorf(q.x), cy = (int)floorf(q.y), cz = (int)floorf(q.z);
    const int x0 = (cx - 1) & (g.DX - 1), x1 = cx & (g.DX - 1), x2 = (cx + 1) & (g.DX - 1);
    const bool contiguous = (x0 + 2 == x2);
#pragma unroll 1
    for (int r = 0; r < 9; ++r) {
        const int dz = r / 3 - 1, dy = r % 3 - 1;
        const int row = (((cz + dz) & (g.DZ - 1)) * g.DY + ((cy + dy) & (g.DY - 1))) * g.DX;
        if (contiguous) {
            unsigned b = __ldg(cell_start + row + x0), e = __ldg(cell_start + row + x2 + 1);
            knn_scan_range(gmap, b, e, gl, q, mine);
        } else {
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

// Solve step shared by the hook kernel and the persistent kernel; executed by ONE thread.
// sums: 28 fp64 totals.  Updates tf (6), state; returns converged.  X_out/AtA_out/AtB_out optional.
__device__ __noinline__ bool lm_solve_dev(int iter, const double* sums, float* tf, LMDeviceState* st, float* scratchA /*36*/, float* scratchV /*36*/,
                                          float* AtA_out, float* AtB_out, float* X_out, int* nsel_out) {
    float AtA[36], AtB[6], X[6];
    int p = 0;
    for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) { float v = (float)sums[p++]; AtA[i * 6 + j] = v; AtA[j * 6 + i] = v; }
    for (int i = 0; i < 6; ++i) AtB[i] = (float)sums[21 + i];
    const int nsel = (int)(sums[27] + 0.5);
    if (nsel_out) *nsel_out = nsel;
    if (AtA_out) for (int i = 0; i < 36; ++i) AtA_out[i] = AtA[i];
    if (AtB_out) for (int i = 0; i < 6; ++i) AtB_out[i] = AtB[i];
    if (X_out) for (int i = 0; i < 6; ++i) X_out[i] = 0.f;
    if (nsel < 50) return false;                                                    // :1178
    qr_solve6(AtA, AtB, X);                                                         // :1240
    if (iter == 0) {                                                                // :1242-1264
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
    if (st->isDegenerate) { float X2[6]; for (int i = 0; i < 6; ++i) X2[i] = X[i]; gemv6(st->matP, X2, X); }   // :1266-1271
    for (int i = 0; i < 6; ++i) tf[i] += X[i];                                      // :1273-1278
    if (X_out) for (int i = 0; i < 6; ++i) X_out[i] = X[i];
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
        bool c = lm_solve_dev(iter, s_sum, tf6, st, s_A, s_V, AtA_out, AtB_out, X_out, nsel_out);
        *conv_out = c ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// scan2MapOptimization: one persistent cooperative kernel runs the whole ≤30-iteration loop.
// ---------------------------------------------------------------------------------------------------------------
struct S2MArgs {
    const float4* scan; Count n_scan;
    const unsigned* cell_start; const float4* gmap; GridDims g; Count m_map;
    float* tf6;                      // in/out transformTobeMapped (device)
    LMDeviceState* st;
    double* partial;                 // [2][gridDim.x][NPROD]
    S2MTrace* trace;
    int max_iters; int force_all;
};

__global__ void __launch_bounds__(S2M_BLOCK) k_scan2map_persistent(S2MArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ float s_tf[6];
    __shared__ float s_t[12];
    __shared__ LMTrig s_trig;
    __shared__ float s_rows[S2M_QPB][8];
    __shared__ double s_red[S2M_BLOCK / 32][LIORF_G][4];
    __shared__ double s_sum[NPROD];
    __shared__ float s_A[36], s_V[36];
    __shared__ LMDeviceState s_st;
    __shared__ int s_conv;

    const int n = a.n_scan.get();
    const int m = a.m_map.get();
    const int gl = threadIdx.x & (LIORF_G - 1);
    const int qslot = threadIdx.x / LIORF_G;
    int pi[4], pj[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) { const int p = gl + LIORF_G * t; pi[t] = prod_i(p < NPROD ? p : NPROD - 1); pj[t] = prod_j(p < NPROD ? p : NPROD - 1); }
    if (threadIdx.x < 6) s_tf[threadIdx.x] = a.tf6[threadIdx.x];
    if (threadIdx.x == 0) s_st = *a.st;
    __syncthreads();
    // guards of scan2MapOptimization (:1297-1300): a map must exist (and hold >= 5 points for the 5-NN), n > 30
    const bool run = (m >= 5) && (n > 30);
    int iters_done = 0, converged = 0;
    if (run) {
        for (int iter = 0; iter < a.max_iters; ++iter) {
            if (threadIdx.x == 0) {
                get_transformation_dev(s_tf[3], s_tf[4], s_tf[5], s_tf[0], s_tf[1], s_tf[2], s_t);
                s_trig = lm_trig(s_tf);
            }
            __syncthreads();
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            for (int q0 = blockIdx.x * S2M_QPB; q0 < n; q0 += gridDim.x * S2M_QPB) {
                const int q = q0 + qslot;
                const bool active = q < n;
                float4 ori = active ? __ldg(a.scan + q) : make_float4(0, 0, 0, 0);
                float4 sel = apply_affine_dev(s_t, ori);
                Top5 nn; knn5_group(sel, a.cell_start, a.gmap, a.g, nn);
                float4 coeff;
                bool f = surf_point_dev(ori, sel, a.gmap, nn, coeff, nullptr) && active;
                float v[8];
                lm_row_dev(s_trig, ori, coeff, v); v[7] = 1.f;
                __syncwarp();
                if (gl < 8) s_rows[qslot][gl] = f ? (gl == 0 ? v[0] : gl == 1 ? v[1] : gl == 2 ? v[2] : gl == 3 ? v[3] : gl == 4 ? v[4] : gl == 5 ? v[5] : gl == 6 ? v[6] : v[7]) : 0.f;
                __syncwarp();
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int p = gl + LIORF_G * t;
                    if (p < NPROD) acc[t] += (double)s_rows[qslot][pi[t]] * (double)s_rows[qslot][pj[t]];
                }
            }
            // block reduction: lanes with equal gl across the 4 groups of a warp, then across warps
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                acc[t] += __shfl_xor_sync(FULL, acc[t], 8);
                acc[t] += __shfl_xor_sync(FULL, acc[t], 16);
            }
            if (lane_id() < LIORF_G) {
#pragma unroll
                for (int t = 0; t < 4; ++t) s_red[warp_id()][gl][t] = acc[t];
            }
            __syncthreads();
            double* part = a.partial + (size_t)(iter & 1) * gridDim.x * NPROD;
            if (threadIdx.x < NPROD) {
                const int p = threadIdx.x, pg = p % LIORF_G, pt = p / LIORF_G;
                double s = 0;
#pragma unroll
                for (int w = 0; w < S2M_BLOCK / 32; ++w) s += s_red[w][pg][pt];
                part[(size_t)blockIdx.x * NPROD + p] = s;
            }
            grid.sync();
            if (threadIdx.x < NPROD) {
                const double* vp = part;
                double s = 0;
                for (int b = 0; b < (int)gridDim.x; ++b) s += __ldcg(vp + (size_t)b * NPROD + threadIdx.x);
                s_sum[threadIdx.x] = s;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                int nsel;
                bool c = lm_solve_dev(iter, s_sum, s_tf, &s_st, s_A, s_V, nullptr, nullptr, nullptr, &nsel);
                s_conv = c ? 1 : 0;
                if (blockIdx.x == 0 && a.trace && iter < S2M_MAX_ITERS) {
                    for (int k = 0; k < 6; ++k) a.trace->pose[iter][k] = s_tf[k];
                    a.trace->nsel[iter] = nsel;
                }
            }
            __syncthreads();
            iters_done = iter + 1;
            if (s_conv) { converged = 1; if (!a.force_all) break; }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int k = 0; k < 6; ++k) a.tf6[k] = s_tf[k];
        *a.st = s_st;
        if (a.trace) { a.trace->iters = iters_done; a.trace->converged = converged; a.trace->degenerate = s_st.isDegenerate; a.trace->ran = run ? 1 : 0; }
    }
}

}  // namespace liorf

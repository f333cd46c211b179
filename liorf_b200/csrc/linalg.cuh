// linalg.cuh — small dense linear algebra kept in registers.
//   colpiv_qr_solve_5x3 : Eigen::ColPivHouseholderQR<Matrix<float,5,3>>::solve  (src/mapOptmization.cpp:1104)
//   qr_solve6 / jacobi6 / lu_invert6 / gemm6 : the OpenCV calls of LMOptimization (src/mapOptmization.cpp:1237-1271)
// Same operation order as the reference's libraries; the file is compiled with -fmad=false.
#pragma once
#include "common.cuh"
#include <cfloat>

namespace liorf {

#define LIORF_SWAPF(a, b) { float _t = (a); (a) = (b); (b) = _t; }

// A: 5 rows of (x,y,z); b = -1 for all rows (matB0.fill(-1)).  x = argmin ||A x - b||.
__device__ __forceinline__ void colpiv_qr_solve_5x3(const float (&Ain)[5][3], float (&x)[3]) {
    float qr[5][3];
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) qr[r][c] = Ain[r][c];
    float hcoef[3], nU[3], nD[3];
    int perm0 = 0, perm1 = 1, perm2 = 2;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < 5; ++r) s += qr[r][c] * qr[r][c];
        nD[c] = nU[c] = sqrtf(s);
    }
    const float eps = FLT_EPSILON;
    float maxn = fmaxf(nU[0], fmaxf(nU[1], nU[2]));
    float th = maxn * eps / 5.0f;
    const float threshold_helper = th * th;
    const float norm_downdate_threshold = sqrtf(eps);
    int nonzero_pivots = 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int big = k; float bign = nU[k];
#pragma unroll
        for (int c = k + 1; c < 3; ++c) if (nU[c] > bign) { bign = nU[c]; big = c; }
        float big_sq = bign * bign;
        if (nonzero_pivots == 3 && big_sq < threshold_helper * (float)(5 - k)) nonzero_pivots = k;
#pragma unroll
        for (int c = k + 1; c < 3; ++c) {
            if (big == c) {
#pragma unroll
                for (int r = 0; r < 5; ++r) LIORF_SWAPF(qr[r][k], qr[r][c]);
                LIORF_SWAPF(nU[k], nU[c]); LIORF_SWAPF(nD[k], nD[c]);
                // indices.swap(k, big)
                int pk = (k == 0) ? perm0 : (k == 1 ? perm1 : perm2);
                int pc = (c == 1) ? perm1 : perm2;
                if (k == 0) perm0 = pc; else if (k == 1) perm1 = pc;
                if (c == 1) perm1 = pk; else perm2 = pk;
            }
        }
        float tailSq = 0.f;
#pragma unroll
        for (int r = k + 1; r < 5; ++r) tailSq += qr[r][k] * qr[r][k];
        float c0 = qr[k][k], beta, tau;
        if (tailSq <= FLT_MIN) {
            tau = 0.f; beta = c0;
#pragma unroll
            for (int r = k + 1; r < 5; ++r) qr[r][k] = 0.f;
        } else {
            beta = sqrtf(c0 * c0 + tailSq);
            if (c0 >= 0.f) beta = -beta;
            float den = c0 - beta;
#pragma unroll
            for (int r = k + 1; r < 5; ++r) qr[r][k] = qr[r][k] / den;
            tau = (beta - c0) / beta;
        }
        hcoef[k] = tau; qr[k][k] = beta;
        if (tau != 0.f) {
#pragma unroll
            for (int c = k + 1; c < 3; ++c) {
                float tmp = 0.f;
#pragma unroll
                for (int r = k + 1; r < 5; ++r) tmp += qr[r][k] * qr[r][c];
                tmp += qr[k][c];
                qr[k][c] -= tau * tmp;
#pragma unroll
                for (int r = k + 1; r < 5; ++r) qr[r][c] -= (tau * qr[r][k]) * tmp;
            }
        }
#pragma unroll
        for (int c = k + 1; c < 3; ++c) {
            if (nU[c] != 0.f) {
                float temp = fabsf(qr[k][c]) / nU[c];
                temp = (1.f + temp) * (1.f - temp);
                temp = temp < 0.f ? 0.f : temp;
                float ratio = nU[c] / nD[c];
                float temp2 = temp * (ratio * ratio);
                if (temp2 <= norm_downdate_threshold) {
                    float s = 0.f;
#pragma unroll
                    for (int r = k + 1; r < 5; ++r) s += qr[r][c] * qr[r][c];
                    nD[c] = sqrtf(s); nU[c] = nD[c];
                } else nU[c] *= sqrtf(temp);
            }
        }
    }
    float cc[5];
#pragma unroll
    for (int r = 0; r < 5; ++r) cc[r] = -1.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float tau = hcoef[k];
        if (k < nonzero_pivots && tau != 0.f) {
            float tmp = 0.f;
#pragma unroll
            for (int r = k + 1; r < 5; ++r) tmp += qr[r][k] * cc[r];
            tmp += cc[k];
            cc[k] -= tau * tmp;
#pragma unroll
            for (int r = k + 1; r < 5; ++r) cc[r] -= (tau * qr[r][k]) * tmp;
        }
    }
    x[0] = x[1] = x[2] = 0.f;
    if (nonzero_pivots == 0) return;
#pragma unroll
    for (int i = 2; i >= 0; --i) {
        if (i < nonzero_pivots) {
            float s = cc[i];
#pragma unroll
            for (int j = i + 1; j < 3; ++j) if (j < nonzero_pivots) s -= qr[i][j] * cc[j];
            cc[i] = s / qr[i][i];
        }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (i < nonzero_pivots) {
            int p = (i == 0) ? perm0 : (i == 1 ? perm1 : perm2);
            if (p == 0) x[0] = cc[i]; else if (p == 1) x[1] = cc[i]; else x[2] = cc[i];
        }
    }
}

// cv::solve(A, b, x, DECOMP_QR) for 6x6 float: hal::QR32f (Householder), fully unrolled → registers.
__device__ __forceinline__ bool qr_solve6(const float* __restrict__ Ain, const float* __restrict__ bin, float* __restrict__ x) {
    float A[36], b[6], vl[6], hF[6];
#pragma unroll
    for (int i = 0; i < 36; ++i) A[i] = Ain[i];
#pragma unroll
    for (int i = 0; i < 6; ++i) b[i] = bin[i];
#pragma unroll
    for (int l = 0; l < 6; ++l) {
        float vlNorm = 0.f;
#pragma unroll
        for (int i = 0; i < 6 - l; ++i) { vl[i] = A[(l + i) * 6 + l]; vlNorm += vl[i] * vl[i]; }
        float tmpV = vl[0];
        vl[0] = vl[0] + (vl[0] < 0.f ? -1.f : 1.f) * sqrtf(vlNorm);
        vlNorm = sqrtf(vlNorm + vl[0] * vl[0] - tmpV * tmpV);
#pragma unroll
        for (int i = 0; i < 6 - l; ++i) vl[i] /= vlNorm;
#pragma unroll
        for (int j = l; j < 6; ++j) {
            float v_lA = 0.f;
#pragma unroll
            for (int i = l; i < 6; ++i) v_lA += vl[i - l] * A[i * 6 + j];
#pragma unroll
            for (int i = l; i < 6; ++i) A[i * 6 + j] -= 2 * vl[i - l] * v_lA;
        }
        hF[l] = vl[0] * vl[0];
#pragma unroll
        for (int i = 1; i < 6 - l; ++i) A[(l + i) * 6 + l] = vl[i] / vl[0];
    }
#pragma unroll
    for (int l = 0; l < 6; ++l) {
        vl[0] = 1.f;
#pragma unroll
        for (int j = 1; j < 6 - l; ++j) vl[j] = A[(j + l) * 6 + l];
        float v_lB = 0.f;
#pragma unroll
        for (int i = l; i < 6; ++i) v_lB += vl[i - l] * b[i];
#pragma unroll
        for (int i = l; i < 6; ++i) b[i] -= 2 * vl[i - l] * v_lB * hF[l];
    }
    const float eps = FLT_EPSILON * 10;
    bool ok = true;
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        if (ok) {
#pragma unroll
            for (int j = 5; j > i; --j) b[i] -= b[j] * A[i * 6 + j];
            if (fabsf(A[i * 6 + i]) < eps) ok = false;
            else b[i] /= A[i * 6 + i];
        }
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) x[i] = ok ? b[i] : 0.f;
    return ok;
}

// qr_solve6 by one warp — same arithmetic, same order, bit-identical results (tests: liorf_debug_qr6 with mode 1).
// Lane j < 6 holds column j of A, lane 6 the right-hand side.  Step l: lane l forms the Householder vector from its column and
// broadcasts it; every column applies the reflector to itself at once, the right-hand side applies the STORED form of the same
// reflector ((1, v1/v0, ..), hF = v0^2 — hal::QR32f transforms b from the stored factors after the sweep; reflector l's stored
// factors never change after step l, so doing it inside the sweep is the same sequence of operations on b).  The six dependent
// column sweeps of the one-thread version become one; the back substitution runs redundantly in every lane, which all return x.
__device__ __forceinline__ bool qr_solve6_warp_cols(float (&col)[6], float* __restrict__ x) {
    const int j = threadIdx.x & 31;
#pragma unroll
    for (int l = 0; l < 6; ++l) {
        float vl[6], vp[6];
        float vlNorm = 0.f;
#pragma unroll
        for (int i = 0; i < 6 - l; ++i) { vl[i] = col[l + i]; vlNorm += vl[i] * vl[i]; }
        float tmpV = vl[0];
        vl[0] = vl[0] + (vl[0] < 0.f ? -1.f : 1.f) * sqrtf(vlNorm);
        vlNorm = sqrtf(vlNorm + vl[0] * vl[0] - tmpV * tmpV);
#pragma unroll
        for (int i = 0; i < 6 - l; ++i) vl[i] /= vlNorm;
#pragma unroll
        for (int i = 0; i < 6 - l; ++i) vl[i] = __shfl_sync(0xffffffffu, vl[i], l);
        const float hF = vl[0] * vl[0];
        vp[0] = 1.f;
#pragma unroll
        for (int i = 1; i < 6 - l; ++i) vp[i] = vl[i] / vl[0];
        float v_lA = 0.f, v_lB = 0.f;
#pragma unroll
        for (int i = l; i < 6; ++i) { v_lA += vl[i - l] * col[i]; v_lB += vp[i - l] * col[i]; }
#pragma unroll
        for (int i = l; i < 6; ++i) {
            const float nA = col[i] - 2 * vl[i - l] * v_lA;
            const float nB = col[i] - 2 * vp[i - l] * v_lB * hF;
            col[i] = j == 6 ? nB : ((j >= l && j < 6) ? nA : col[i]);
        }
#pragma unroll
        for (int i = 1; i < 6 - l; ++i) if (j == l) col[l + i] = vp[i];
    }
    float b[6], R[6][6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        b[i] = __shfl_sync(0xffffffffu, col[i], 6);
#pragma unroll
        for (int jj = i; jj < 6; ++jj) R[i][jj] = __shfl_sync(0xffffffffu, col[i], jj);
    }
    const float eps = FLT_EPSILON * 10;
    bool ok = true;
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        if (ok) {
#pragma unroll
            for (int jj = 5; jj > i; --jj) b[i] -= b[jj] * R[i][jj];
            if (fabsf(R[i][i]) < eps) ok = false;
            else b[i] /= R[i][i];
        }
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) x[i] = ok ? b[i] : 0.f;
    return ok;
}
__device__ __forceinline__ bool qr_solve6_warp(const float* __restrict__ Ain, const float* __restrict__ bin, float* __restrict__ x) {
    const int j = threadIdx.x & 31;
    float col[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) col[i] = j < 6 ? Ain[i * 6 + j] : (j == 6 ? bin[i] : 0.f);
    return qr_solve6_warp_cols(col, x);
}

__device__ __forceinline__ float cv_hypot(float a, float b) {
    a = fabsf(a); b = fabsf(b);
    if (a > b) { b /= a; return a * sqrtf(1 + b * b); }
    if (b > 0) { a /= b; return b * sqrtf(1 + a * a); }
    return 0.f;
}

// cv::eigen for symmetric 6x6 float (JacobiImpl_): W descending, eigenvectors in ROWS of V.
// A, V are caller-provided 36-float scratch (shared or local memory; indices are data dependent).
__device__ __noinline__ void jacobi6(float* A, float* W, float* V) {
    const int n = 6; const float eps = FLT_EPSILON;
    int indR[6], indC[6]; int i, j, k, m; float mv;
    for (i = 0; i < n; ++i) { for (j = 0; j < n; ++j) V[i * 6 + j] = 0.f; V[i * 6 + i] = 1.f; }
    for (k = 0; k < n; ++k) {
        W[k] = A[7 * k];
        if (k < n - 1) {
            for (m = k + 1, mv = fabsf(A[6 * k + m]), i = k + 2; i < n; ++i) { float val = fabsf(A[6 * k + i]); if (mv < val) mv = val, m = i; }
            indR[k] = m;
        }
        if (k > 0) {
            for (m = 0, mv = fabsf(A[k]), i = 1; i < k; ++i) { float val = fabsf(A[6 * i + k]); if (mv < val) mv = val, m = i; }
            indC[k] = m;
        }
    }
    const int maxIters = n * n * 30;
    for (int iters = 0; iters < maxIters; ++iters) {
        for (k = 0, mv = fabsf(A[indR[0]]), i = 1; i < n - 1; ++i) { float val = fabsf(A[6 * i + indR[i]]); if (mv < val) mv = val, k = i; }
        int l = indR[k];
        for (i = 1; i < n; ++i) { float val = fabsf(A[6 * indC[i] + i]); if (mv < val) mv = val, k = indC[i], l = i; }
        float p = A[6 * k + l];
        if (fabsf(p) <= eps) break;
        float y = (float)((double)(W[l] - W[k]) * 0.5);
        float t = fabsf(y) + cv_hypot(p, y);
        float s = cv_hypot(p, t);
        float c = t / s;
        s = p / s; t = (p / t) * p;
        if (y < 0) s = -s, t = -t;
        A[6 * k + l] = 0;
        W[k] -= t; W[l] += t;
        float a0, b0;
#define LIORF_ROT(v0, v1) a0 = v0, b0 = v1, v0 = a0 * c - b0 * s, v1 = a0 * s + b0 * c
        for (i = 0; i < k; ++i) LIORF_ROT(A[6 * i + k], A[6 * i + l]);
        for (i = k + 1; i < l; ++i) LIORF_ROT(A[6 * k + i], A[6 * i + l]);
        for (i = l + 1; i < n; ++i) LIORF_ROT(A[6 * k + i], A[6 * l + i]);
        for (i = 0; i < n; ++i) LIORF_ROT(V[6 * k + i], V[6 * l + i]);
#undef LIORF_ROT
        for (j = 0; j < 2; ++j) {
            int idx = j == 0 ? k : l;
            if (idx < n - 1) {
                for (m = idx + 1, mv = fabsf(A[6 * idx + m]), i = idx + 2; i < n; ++i) { float val = fabsf(A[6 * idx + i]); if (mv < val) mv = val, m = i; }
                indR[idx] = m;
            }
            if (idx > 0) {
                for (m = 0, mv = fabsf(A[idx]), i = 1; i < idx; ++i) { float val = fabsf(A[6 * i + idx]); if (mv < val) mv = val, m = i; }
                indC[idx] = m;
            }
        }
    }
    for (k = 0; k < n - 1; ++k) {
        m = k;
        for (i = k + 1; i < n; ++i) if (W[m] < W[i]) m = i;
        if (k != m) { LIORF_SWAPF(W[m], W[k]); for (i = 0; i < n; ++i) LIORF_SWAPF(V[6 * m + i], V[6 * k + i]); }
    }
}

// cv::Mat::inv() (DECOMP_LU) for 6x6 float: hal::LU32f on [A | I].  A is destroyed.
__device__ __noinline__ bool lu_invert6(float* A, float* inv) {
    const int m = 6;
    for (int i = 0; i < 36; ++i) inv[i] = 0.f;
    for (int i = 0; i < 6; ++i) inv[i * 6 + i] = 1.f;
    const float eps = FLT_EPSILON * 10;
    for (int i = 0; i < m; ++i) {
        int k = i;
        for (int j = i + 1; j < m; ++j) if (fabsf(A[j * 6 + i]) > fabsf(A[k * 6 + i])) k = j;
        if (fabsf(A[k * 6 + i]) < eps) { for (int q = 0; q < 36; ++q) inv[q] = 0.f; return false; }
        if (k != i) {
            for (int j = i; j < m; ++j) LIORF_SWAPF(A[i * 6 + j], A[k * 6 + j]);
            for (int j = 0; j < m; ++j) LIORF_SWAPF(inv[i * 6 + j], inv[k * 6 + j]);
        }
        float d = -1 / A[i * 6 + i];
        for (int j = i + 1; j < m; ++j) {
            float alpha = A[j * 6 + i] * d;
            for (int q = i + 1; q < m; ++q) A[j * 6 + q] += alpha * A[i * 6 + q];
            for (int q = 0; q < m; ++q) inv[j * 6 + q] += alpha * inv[i * 6 + q];
        }
    }
    for (int i = m - 1; i >= 0; --i)
        for (int j = 0; j < m; ++j) {
            float s = inv[i * 6 + j];
            for (int k = i + 1; k < m; ++k) s -= A[i * 6 + k] * inv[k * 6 + j];
            inv[i * 6 + j] = s / A[i * 6 + i];
        }
    return true;
}

// cv::gemm small path: double accumulator, rounded to float
__device__ __forceinline__ void gemm6(const float* A, const float* B, float* C) {
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) {
        double s = 0; for (int k = 0; k < 6; ++k) s += (double)A[i * 6 + k] * (double)B[k * 6 + j];
        C[i * 6 + j] = (float)s;
    }
}
__device__ __forceinline__ void gemv6(const float* A, const float* x, float* y) {
#pragma unroll
    for (int i = 0; i < 6; ++i) { double s = 0;
#pragma unroll
        for (int k = 0; k < 6; ++k) s += (double)A[i * 6 + k] * (double)x[k];
        y[i] = (float)s; }
}

}  // namespace liorf

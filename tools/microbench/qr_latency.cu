// Latency microbenchmark behind the reducer's 6x6 solve: clock64 around (a) chains of dependent fp32 adds / IEEE divisions / square roots,
// (b) six independent IEEE divisions, (c) the one-thread and the one-warp Householder solves of csrc/linalg.cuh.
// Measured on B200 (profiles/r02_qr_latency.txt): dependent FADD 8 cycles, IEEE division 44, square root + add 47; six independent IEEE
// divisions 582 (97 each: the per-operation slow-path guard serialises them); one-thread solve 4.1-4.3k cycles, one-warp solve 5.3-5.7k in
// isolation.  Writing nvcc's fast paths out so that a batch shares one range check gave bit-identical results (268 M operands, 20 000
// systems) but a SLOWER solve (5.3k one thread, 6.2k in the solver): dropped.
// build + run on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -std=c++17 --expt-relaxed-constexpr -o /tmp/qrl tools/microbench/qr_latency.cu && /tmp/qrl
#include <cstdio>
#include "../../liorf_b200/csrc/linalg.cuh"
using namespace liorf;
__global__ void k(const float* A, const float* b, float* out, long long* t) {
    float x = A[0], y = A[7], acc = 0.f;
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) acc = acc + x;                      // 64 dependent FADD
    long long t1 = clock64();
    float d = y;
#pragma unroll
    for (int i = 0; i < 16; ++i) d = x / d;                          // 16 dependent IEEE divisions
    long long t2 = clock64();
    float s = y;
#pragma unroll
    for (int i = 0; i < 16; ++i) s = sqrtf(s + x);                   // 16 dependent IEEE square roots (+ add)
    long long t3 = clock64();
    float q[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) q[i] = A[i + 1] / y;                 // 6 independent IEEE divisions
    float qs = q[0] + q[1] + q[2] + q[3] + q[4] + q[5];
    long long t4 = clock64();
    float qs2 = 0.f;
    long long t5 = clock64();
    float xs[6], xw[6];
    qr_solve6(A, b, xs);
    long long t6 = clock64();
    qr_solve6_warp(A, b, xw);
    long long t7 = clock64();
    long long t8 = clock64();
    // fp64 sin / cos of a float angle rounded to float (the six trig values of an LM iteration), and dependent fp64 FMAs
    const float ang = A[2] * 0.01f;
    const float sn = (float)sin((double)ang), cs = (float)cos((double)ang);
    long long t9 = clock64();
    double dd = 0.0;
    long long t10 = t9;
    if (threadIdx.x == 0) {
        out[0] = acc + d + s + qs + qs2 + sn + cs + (float)dd; for (int i = 0; i < 6; ++i) out[1 + i] = xs[i] + xw[i];
        t[0] = t1 - t0; t[1] = t2 - t1; t[2] = t3 - t2; t[3] = t4 - t3; t[4] = t5 - t4; t[5] = t6 - t5; t[6] = t7 - t6; t[7] = t8 - t7; t[9] = t9 - t8; t[10] = t10 - t9; t[8] = 0; for (int i = 0; i < 6; ++i) t[8] += (__float_as_uint(xs[i]) != __float_as_uint(xw[i]));
    }
}
int main() {
    float hA[36], hb[6];
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) hA[i * 6 + j] = (i == j ? 50.f : 0.f) + 1.f / (1 + i + j);
    for (int i = 0; i < 6; ++i) hb[i] = 1.f + i;
    float *dA, *db, *dout; long long* dt;
    cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&db, sizeof(hb)); cudaMalloc(&dout, 64); cudaMalloc(&dt, 128);
    cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice); cudaMemcpy(db, hb, sizeof(hb), cudaMemcpyHostToDevice);
    long long ht[16];
    for (int rep = 0; rep < 3; ++rep) {
        k<<<1, 32>>>(dA, db, dout, dt);
        cudaMemcpy(ht, dt, 88, cudaMemcpyDeviceToHost);
        printf("cycles: 64 dependent FADD %lld | 16 dependent IEEE div %lld | 16 dependent IEEE sqrt+add %lld | 6 independent IEEE div %lld | (unused %lld) | one-thread 6x6 QR solve %lld | one-warp 6x6 QR solve %lld | (unused %lld) (result words differing between the two solves: %lld) | (float)sin + (float)cos of a double %lld (%lld)\n",
               ht[0], ht[1], ht[2], ht[3], ht[4], ht[5], ht[6], ht[7], ht[8], ht[9], ht[10]);
    }
    return 0;
}

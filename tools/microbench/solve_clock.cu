// microbenchmark: cycles of the pieces of the LM solve step run by ONE thread (as the last CTA of the persistent solver does)
#include "../../liorf_b200/csrc/scan2map.cuh"
using namespace liorf;
__global__ void k(const double* sums_g, long long* out, float* sink) {
    __shared__ double s_sum[NPROD]; __shared__ float s_A[36], s_V[36]; __shared__ LMDeviceState s_st;
    if (threadIdx.x < NPROD) s_sum[threadIdx.x] = sums_g[threadIdx.x];
    __syncthreads();
    if (threadIdx.x != 0) return;
    for (int rep = 0; rep < 3; ++rep) {
        long long t0 = clock64();
        float AtA[36], AtB[6], X[6];
#pragma unroll
        for (int p = 0; p < 21; ++p) { const float v = (float)s_sum[p]; AtA[prod_i(p) * 6 + prod_j(p)] = v; AtA[prod_j(p) * 6 + prod_i(p)] = v; }
#pragma unroll
        for (int i = 0; i < 6; ++i) AtB[i] = (float)s_sum[21 + i];
        long long t1 = clock64();
        qr_solve6(AtA, AtB, X);
        long long t2 = clock64();
        bool cert = certify_non_degenerate(AtA);
        long long t3 = clock64();
        const float r2d = 57.29578f;
        double a0 = (double)(X[0] * r2d), a1 = (double)(X[1] * r2d), a2 = (double)(X[2] * r2d);
        double q0 = (double)(X[3] * 100), q1 = (double)(X[4] * 100), q2 = (double)(X[5] * 100);
        float deltaR = (float)sqrt(a0 * a0 + a1 * a1 + a2 * a2);
        float deltaT = (float)sqrt(q0 * q0 + q1 * q1 + q2 * q2);
        bool conv = (double)deltaR < 0.05 && (double)deltaT < 0.05;
        long long t4 = clock64();
        float tf[6] = {0, 0, 0, 0, 0, 0}; int nsel;
        bool c2 = lm_solve_dev<true>(1, s_sum, tf, &s_st, s_A, s_V, nullptr, nullptr, nullptr, &nsel);
        long long t5 = clock64();
        bool c3 = lm_solve_dev<true>(0, s_sum, tf, &s_st, s_A, s_V, nullptr, nullptr, nullptr, &nsel);
        long long t6 = clock64();
        out[rep * 8 + 0] = t1 - t0; out[rep * 8 + 1] = t2 - t1; out[rep * 8 + 2] = t3 - t2; out[rep * 8 + 3] = t4 - t3; out[rep * 8 + 4] = t5 - t4; out[rep * 8 + 5] = t6 - t5;
        sink[rep] = X[0] + (cert ? 1.f : 0.f) + (conv ? 1.f : 0.f) + tf[0] + (c2 ? 1.f : 0.f) + (c3 ? 1.f : 0.f);
    }
}
int main() {
    double h[NPROD]; for (int p = 0; p < NPROD; ++p) h[p] = 0;
    // a well conditioned SPD system: AtA = diag-dominant
    for (int p = 0; p < 21; ++p) h[p] = (prod_i(p) == prod_j(p)) ? 5000.0 + 300 * prod_i(p) : 37.0 * (p % 5) - 60;
    for (int i = 0; i < 6; ++i) h[21 + i] = 3.0 + i; h[27] = 3000;
    double* d; long long* o; float* s; cudaMalloc(&d, sizeof(h)); cudaMalloc(&o, 24 * 8); cudaMalloc(&s, 16);
    cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
    k<<<1, 64>>>(d, o, s);
    long long ho[24]; cudaMemcpy(ho, o, sizeof(ho), cudaMemcpyDeviceToHost);
    const char* nm[6] = {"sums->AtA/AtB (34 F2F)", "qr_solve6", "certificate", "convergence test", "lm_solve_dev iter>0", "lm_solve_dev iter 0"};
    for (int rep = 0; rep < 3; ++rep) { printf("rep %d:", rep); for (int k2 = 0; k2 < 6; ++k2) printf("  %s %lld cyc;", nm[k2], ho[rep * 8 + k2]); printf("\n"); }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

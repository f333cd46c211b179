// common.cuh — shared types and helpers of the B200 (sm_100a) scan-to-map library.
// The whole library is compiled with -fmad=false: parity-critical fp32 expressions keep the
// reference's operation order (x86-64 baseline build, no FMA: CMakeLists.txt:7).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cmath>

#define LIORF_OK 0
#define LIORF_ERR_CUDA -1
#define LIORF_ERR_ARG -2
#define LIORF_ERR_STATE -3
#define LIORF_ERR_DEVICE_FLAG -4

#define CUDA_TRY(expr)                                                                       \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            fprintf(stderr, "[liorf_b200] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(_e), __FILE__, __LINE__, \
                    cudaGetErrorString(_e));                                                 \
            return LIORF_ERR_CUDA;                                                           \
        }                                                                                    \
    } while (0)

namespace liorf {

constexpr int kNumSMs = 148;          // B200: 2 dies x 74 SMs; grids are sized in multiples of this
constexpr unsigned FULL = 0xffffffffu;

// growable device buffer (capacity only grows; contents are NOT preserved across growth unless asked)
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    int reserve(size_t n, cudaStream_t s = nullptr, bool keep = false) {
        if (n <= cap) return LIORF_OK;
        size_t ncap = cap ? cap : 1024;
        while (ncap < n) ncap = 2 * ncap + 1024;          // doubling: growth (a cudaMalloc + cudaFree sync) must stay rare
        T* np_ = nullptr;
        CUDA_TRY(cudaMalloc(&np_, ncap * sizeof(T)));
        if (keep && p && cap) CUDA_TRY(cudaMemcpyAsync(np_, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, s));
        if (p) { if (keep) CUDA_TRY(cudaStreamSynchronize(s)); CUDA_TRY(cudaFree(p)); }
        p = np_; cap = ncap;
        return LIORF_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// correctly-rounded-in-practice float trig: evaluate in fp64, round once (see DESIGN.md "transcendentals")
__device__ __forceinline__ float sin_f(float x) { return (float)sin((double)x); }
__device__ __forceinline__ float cos_f(float x) { return (float)cos((double)x); }

// pcl::getTransformation (float), row-major 3x4 — mirrors oracle get_transformation
__device__ __forceinline__ void get_transformation_dev(float x, float y, float z, float roll, float pitch, float yaw, float t[12]) {
    float A = cos_f(yaw), B = sin_f(yaw), C = cos_f(pitch), D = sin_f(pitch);
    float E = cos_f(roll), F = sin_f(roll), DE = D * E, DF = D * F;
    t[0] = A * C;  t[1] = A * DF - B * E;  t[2]  = B * F + A * DE;  t[3]  = x;
    t[4] = B * C;  t[5] = A * E + B * DF;  t[6]  = B * DE - A * F;  t[7]  = y;
    t[8] = -D;     t[9] = C * F;           t[10] = C * E;           t[11] = z;
}

// pointAssociateToMap op order (src/mapOptmization.cpp:304-306)
__device__ __forceinline__ float4 apply_affine_dev(const float* __restrict__ t, float4 p) {
    float4 o;
    o.x = t[0] * p.x + t[1] * p.y + t[2]  * p.z + t[3];
    o.y = t[4] * p.x + t[5] * p.y + t[6]  * p.z + t[7];
    o.z = t[8] * p.x + t[9] * p.y + t[10] * p.z + t[11];
    o.w = p.w;
    return o;
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

}  // namespace liorf

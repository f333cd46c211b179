// voxelgrid.cuh — pcl::VoxelGrid<PointXYZI> leaf-centroid downsample on the GPU.
// Replaces the filter behind downsampleCurrentScan (src/mapOptmization.cpp:1061-1067) and
// downSizeFilterLocalMapSurf in extractCloud (src/mapOptmization.cpp:1037-1038).
//
// Pipeline (all counts stay on the device):
//   k_vg_minmax   : componentwise min/max, last block derives min_b / div_b / multipliers and PCL's overflow guard
//   k_vg_keys     : int voxel index per point (bit-exact with PCL's float floor arithmetic)
//   radix sort    : stable (key, point index) sort — 4 x 8-bit passes
//   scan          : segment heads → output slot; segment start table
//   k_vg_centroid : one thread per voxel; sequential fp32 sums in ascending input index, divided by (float)count
// HBM traffic (algorithmic): 16 N in + 16 N_out out; the sort moves 8 N per pass on top (L2 resident for one scan).
#pragma once
#include "prims.cuh"

namespace liorf {

struct VoxMeta {
    float minp[3], maxp[3];
    int min_b[3], div_b[3];
    int mul1, mul2;
    float inv;
    int overflow;     // PCL: "Leaf size is too small for the input dataset" → output = input
    int n;
};

struct VoxelGridWork {
    DevBuf<float> partial;          // per-block min/max (6 floats each)
    DevBuf<unsigned> keys;
    DevBuf<unsigned> seg_start;     // start of each voxel segment in the sorted order (+ sentinel)
    VoxMeta* meta = nullptr;        // device
    int* mm_counter = nullptr;      // device, zero-initialised
    SortWork sort;
    ScanWork scan;
};

constexpr int VG_MM_BLOCK = 256;

__global__ void __launch_bounds__(VG_MM_BLOCK) k_vg_minmax(const float4* __restrict__ pts, Count cnt, float leaf, float* __restrict__ partial,
                                                           int* counter, VoxMeta* __restrict__ meta) {
    const int n = cnt.get();
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 p = pts[i];
        mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
        mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
        mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
    }
    __shared__ float s_red[VG_MM_BLOCK / 32][6];
    __shared__ bool s_last;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { mn[a] = fminf(mn[a], __shfl_xor_sync(FULL, mn[a], o)); mx[a] = fmaxf(mx[a], __shfl_xor_sync(FULL, mx[a], o)); }
    if (lane_id() == 0) { for (int a = 0; a < 3; ++a) { s_red[warp_id()][a] = mn[a]; s_red[warp_id()][3 + a] = mx[a]; } }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < VG_MM_BLOCK / 32; ++w) for (int a = 0; a < 3; ++a) { mn[a] = fminf(mn[a], s_red[w][a]); mx[a] = fmaxf(mx[a], s_red[w][3 + a]); }
        for (int a = 0; a < 3; ++a) { partial[blockIdx.x * 6 + a] = mn[a]; partial[blockIdx.x * 6 + 3 + a] = mx[a]; }
        __threadfence();
        int t = atomicAdd(counter, 1);
        s_last = (t == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x < 32) {                      // last block: warp 0 folds the per-block partials (min/max are order independent)
        __threadfence();
        for (int a = 0; a < 3; ++a) { mn[a] = INFINITY; mx[a] = -INFINITY; }
        for (int b = threadIdx.x; b < (int)gridDim.x; b += 32)
            for (int a = 0; a < 3; ++a) { mn[a] = fminf(mn[a], __ldcg(partial + b * 6 + a)); mx[a] = fmaxf(mx[a], __ldcg(partial + b * 6 + 3 + a)); }
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { mn[a] = fminf(mn[a], __shfl_xor_sync(FULL, mn[a], o)); mx[a] = fmaxf(mx[a], __shfl_xor_sync(FULL, mx[a], o)); }
    }
    if (threadIdx.x == 0) {
        *counter = 0;
        VoxMeta m;
        m.n = n;
        m.inv = 1.0f / leaf;
        if (n == 0) { for (int a = 0; a < 3; ++a) { mn[a] = 0.f; mx[a] = 0.f; } }
        long long d[3];
        for (int a = 0; a < 3; ++a) {
            m.minp[a] = mn[a]; m.maxp[a] = mx[a];
            d[a] = (long long)((mx[a] - mn[a]) * m.inv) + 1;
            m.min_b[a] = (int)floorf(mn[a] * m.inv);
            int max_b = (int)floorf(mx[a] * m.inv);
            m.div_b[a] = max_b - m.min_b[a] + 1;
        }
        m.overflow = (n > 0 && d[0] * d[1] * d[2] > 2147483647LL) ? 1 : 0;
        m.mul1 = m.div_b[0]; m.mul2 = m.div_b[0] * m.div_b[1];
        *meta = m;
    }
}

__global__ void __launch_bounds__(256) k_vg_keys(const float4* __restrict__ pts, Count cnt, const VoxMeta* __restrict__ meta,
                                                unsigned* __restrict__ keys) {
    const int n = cnt.get();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float inv = meta->inv;
    if (meta->overflow) { keys[i] = (unsigned)i; return; }           // every point its own voxel ⇒ output == input
    float4 p = pts[i];
    int i0 = (int)(floorf(p.x * inv) - (float)meta->min_b[0]);
    int i1 = (int)(floorf(p.y * inv) - (float)meta->min_b[1]);
    int i2 = (int)(floorf(p.z * inv) - (float)meta->min_b[2]);
    keys[i] = (unsigned)(i0 + i1 * meta->mul1 + i2 * meta->mul2);
}

struct VgHeadLoad {
    const unsigned* keys;
    __device__ __forceinline__ unsigned operator()(int i) const { return (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u; }
};
struct VgHeadStore {
    unsigned* seg_start; Count cnt;
    __device__ __forceinline__ void operator()(int i, unsigned v, unsigned excl) const {
        if (v) seg_start[excl] = (unsigned)i;
        if (i == cnt.get() - 1) seg_start[excl + v] = (unsigned)(i + 1);       // sentinel
    }
};

__global__ void __launch_bounds__(128) k_vg_centroid(const float4* __restrict__ pts, const unsigned* __restrict__ sorted_idx,
                                                     const unsigned* __restrict__ seg_start, const unsigned* __restrict__ n_out,
                                                     float4* __restrict__ out, int* __restrict__ membership,
                                                     const unsigned* __restrict__ sorted_keys, int* __restrict__ out_keys) {
    const int nseg = (int)*n_out;
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    unsigned b = seg_start[s], e = seg_start[s + 1];
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    unsigned k = b;
    for (; k + 4 <= e; k += 4) {                  // gathers issued four at a time; the adds stay in ascending input index
        unsigned i0 = sorted_idx[k], i1 = sorted_idx[k + 1], i2 = sorted_idx[k + 2], i3 = sorted_idx[k + 3];
        float4 p0 = __ldg(pts + i0), p1 = __ldg(pts + i1), p2 = __ldg(pts + i2), p3 = __ldg(pts + i3);
        sx += p0.x; sy += p0.y; sz += p0.z; si += p0.w;
        sx += p1.x; sy += p1.y; sz += p1.z; si += p1.w;
        sx += p2.x; sy += p2.y; sz += p2.z; si += p2.w;
        sx += p3.x; sy += p3.y; sz += p3.z; si += p3.w;
        if (membership) { membership[i0] = s; membership[i1] = s; membership[i2] = s; membership[i3] = s; }
    }
    for (; k < e; ++k) {
        unsigned idx = sorted_idx[k];
        float4 p = __ldg(pts + idx);
        sx += p.x; sy += p.y; sz += p.z; si += p.w;
        if (membership) membership[idx] = s;
    }
    float c = (float)(e - b);
    out[s] = make_float4(sx / c, sy / c, sz / c, si / c);
    if (out_keys) out_keys[s] = (int)sorted_keys[b];
}

// out must hold cnt.bound points.  n_out_dev receives the voxel count (device int).
inline int voxel_grid_device(const float4* in, Count cnt, float leaf, float4* out, int* n_out_dev, int* membership, int* out_keys,
                             VoxelGridWork& w, cudaStream_t s) {
    const int nb = cnt.bound;
    if (nb <= 0) { CUDA_TRY(cudaMemsetAsync(n_out_dev, 0, sizeof(int), s)); return LIORF_OK; }
    int rc;
    int mmb = (nb + VG_MM_BLOCK * 8 - 1) / (VG_MM_BLOCK * 8); if (mmb > kNumSMs) mmb = kNumSMs;
    if ((rc = w.partial.reserve((size_t)mmb * 6))) return rc;
    if ((rc = w.keys.reserve(nb))) return rc;
    if ((rc = w.seg_start.reserve((size_t)nb + 1))) return rc;
    k_vg_minmax<<<mmb, VG_MM_BLOCK, 0, s>>>(in, cnt, leaf, w.partial.p, w.mm_counter, w.meta);
    k_vg_keys<<<(nb + 255) / 256, 256, 0, s>>>(in, cnt, w.meta, w.keys.p);
    unsigned* sorted_idx = nullptr;
    if ((rc = radix_sort_pairs_iota(w.keys.p, cnt, w.sort, &sorted_idx, s))) return rc;
    if ((rc = launch_scan(cnt, VgHeadLoad{w.keys.p}, VgHeadStore{w.seg_start.p, cnt}, w.scan, (unsigned*)n_out_dev, s))) return rc;
    k_vg_centroid<<<(nb + 127) / 128, 128, 0, s>>>(in, sorted_idx, w.seg_start.p, (const unsigned*)n_out_dev, out, membership, w.keys.p, out_keys);
    CUDA_TRY(cudaGetLastError());
    return LIORF_OK;
}

}  // namespace liorf

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
    bool small_attr_set = false;
    bool force_large = false;       // tests: run the multi-kernel path on small clouds too
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

// ---------------------------------------------------------------------------------------------------------------
// Tiny clouds (<= 3072 points: Livox scans after point_filter_num, key-pose clouds, test fixtures): the whole filter in ONE
// launch by ONE CTA instead of nine dependent launches (22 us instead of 61 us at 3000 points).  Measured on B200 the
// single CTA stops paying at ~10 k points — one SM cannot issue the sort's instructions fast enough — so larger clouds
// keep the multi-kernel radix path.
//   min/max → keys → stable LSD radix sort of (key, input index) in shared memory (4-bit digits, 4 items per thread held in
//   registers in blocked order, ranks from a block-wide scan of per-thread digit counts — only as many passes as the voxel index
//   has bits) → segment heads → exclusive scan → centroids.  Results are bit-identical to the large path.
// ---------------------------------------------------------------------------------------------------------------
constexpr int VGS_THREADS = 768;
constexpr int VGS_E = 4;                              // items per thread, held in registers in blocked order
constexpr int VGS_CAP = VGS_E * VGS_THREADS;          // 3072 points: (key u32 + index u16) x 3072 = 18 KB of shared memory
constexpr int VGS_SMEM = VGS_CAP * 6 + 16 * VGS_THREADS * 2;     // keys + indices + digit counters


// stable block-wide LSD radix sort of P = E * VGS_THREADS (key, index) pairs living in shared memory.  Items are held in
// registers in blocked order (indices packed two per register, in-thread ranks four per register).
template <int E>
__device__ __forceinline__ void vgs_radix_sort(unsigned* s_key, unsigned short* s_idx, unsigned short* s_cnt, unsigned* s_wsum, int bits) {
    const int tid = threadIdx.x;
    unsigned key[E]; unsigned ip[E / 2];
    const unsigned* s_idx2 = reinterpret_cast<const unsigned*>(s_idx);
#pragma unroll
    for (int u = 0; u < E; ++u) key[u] = s_key[tid * E + u];
#pragma unroll
    for (int u = 0; u < E / 2; ++u) ip[u] = s_idx2[tid * (E / 2) + u];
    for (int shift = 0; shift < bits; shift += 4) {
        unsigned long long lo = 0, hi = 0;                      // sixteen 8-bit counters (E <= 32 fits)
        unsigned rk[E / 4];
#pragma unroll
        for (int u = 0; u < E; ++u) {
            const unsigned d = (key[u] >> shift) & 15u;
            const unsigned sh = (d & 7u) * 8u;
            const unsigned long long sel = d < 8u ? lo : hi;
            const unsigned r = (unsigned)(sel >> sh) & 0xffu;
            if ((u & 3) == 0) rk[u / 4] = r; else rk[u / 4] |= r << (8 * (u & 3));
            const unsigned long long inc = 1ull << sh;
            lo += d < 8u ? inc : 0ull; hi += d < 8u ? 0ull : inc;
        }
#pragma unroll
        for (int d = 0; d < 16; ++d) s_cnt[d * VGS_THREADS + tid] = (unsigned short)(((d < 8 ? lo : hi) >> ((d & 7) * 8)) & 0xffu);
        __syncthreads();
        // exclusive scan over the 16 x VGS_THREADS counters in (digit, thread) order: thread t rakes entries [16 t, 16 t + 16)
        unsigned sum = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) sum += s_cnt[tid * 16 + k];
        unsigned incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(FULL, incl, o); if (lane_id() >= o) incl += v; }
        if (lane_id() == 31) s_wsum[warp_id()] = incl;
        __syncthreads();
        if (warp_id() == 0) {
            const unsigned v = lane_id() < VGS_THREADS / 32 ? s_wsum[lane_id()] : 0u; unsigned inc2 = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned w = __shfl_up_sync(FULL, inc2, o); if (lane_id() >= o) inc2 += w; }
            if (lane_id() < VGS_THREADS / 32) s_wsum[lane_id()] = inc2 - v;
        }
        __syncthreads();
        unsigned run = s_wsum[warp_id()] + incl - sum;
#pragma unroll
        for (int k = 0; k < 16; ++k) { const unsigned v = s_cnt[tid * 16 + k]; s_cnt[tid * 16 + k] = (unsigned short)run; run += v; }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < E; ++u) {
            const unsigned d = (key[u] >> shift) & 15u;
            const unsigned dest = (unsigned)s_cnt[d * VGS_THREADS + tid] + ((rk[u / 4] >> (8 * (u & 3))) & 0xffu);
            s_key[dest] = key[u]; s_idx[dest] = (unsigned short)((u & 1) ? (ip[u / 2] >> 16) : (ip[u / 2] & 0xffffu));
        }
        __syncthreads();
        if (shift + 4 < bits) {
#pragma unroll
            for (int u = 0; u < E; ++u) key[u] = s_key[tid * E + u];
#pragma unroll
            for (int u = 0; u < E / 2; ++u) ip[u] = s_idx2[tid * (E / 2) + u];
        }
    }
}

__global__ void __launch_bounds__(VGS_THREADS, 1) k_vg_small(const float4* __restrict__ pts, Count cnt, float leaf, float4* __restrict__ out, int* __restrict__ n_out_dev,
                                                            int* __restrict__ membership, int* __restrict__ out_keys, VoxMeta* __restrict__ meta_out,
                                                            unsigned* __restrict__ seg_start) {
    extern __shared__ unsigned char vgs_smem[];
    unsigned* s_key = reinterpret_cast<unsigned*>(vgs_smem);
    unsigned short* s_idx = reinterpret_cast<unsigned short*>(vgs_smem + (size_t)VGS_CAP * 4);
    unsigned short* s_cnt = reinterpret_cast<unsigned short*>(vgs_smem + (size_t)VGS_CAP * 6);
    __shared__ float s_red[VGS_THREADS / 32][6];
    __shared__ VoxMeta s_meta;
    __shared__ unsigned s_wsum[32];
    __shared__ unsigned s_total;
    const int n = cnt.get();
    const int tid = threadIdx.x;
    if (n <= 0) { if (tid == 0) { *n_out_dev = 0; } return; }
    // ---- 1. bounds (min/max are order independent) ----
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = tid; i < n; i += VGS_THREADS) {
        const float4 p = pts[i];
        mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
        mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
        mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { mn[a] = fminf(mn[a], __shfl_xor_sync(FULL, mn[a], o)); mx[a] = fmaxf(mx[a], __shfl_xor_sync(FULL, mx[a], o)); }
    if (lane_id() == 0) { for (int a = 0; a < 3; ++a) { s_red[warp_id()][a] = mn[a]; s_red[warp_id()][3 + a] = mx[a]; } }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < VGS_THREADS / 32; ++w) for (int a = 0; a < 3; ++a) { mn[a] = fminf(mn[a], s_red[w][a]); mx[a] = fmaxf(mx[a], s_red[w][3 + a]); }
        VoxMeta m;
        m.n = n;
        m.inv = 1.0f / leaf;
        long long d[3];
        for (int a = 0; a < 3; ++a) {
            m.minp[a] = mn[a]; m.maxp[a] = mx[a];
            d[a] = (long long)((mx[a] - mn[a]) * m.inv) + 1;
            m.min_b[a] = (int)floorf(mn[a] * m.inv);
            const int max_b = (int)floorf(mx[a] * m.inv);
            m.div_b[a] = max_b - m.min_b[a] + 1;
        }
        m.overflow = (d[0] * d[1] * d[2] > 2147483647LL) ? 1 : 0;
        m.mul1 = m.div_b[0]; m.mul2 = m.div_b[0] * m.div_b[1];
        s_meta = m; *meta_out = m;
    }
    __syncthreads();
    // ---- 2. keys (same float arithmetic as k_vg_keys), padded to VGS_CAP items with keys that sort last ----
    const int P = VGS_CAP;
    int bits;
    {
        const float inv = s_meta.inv; const int b0 = s_meta.min_b[0], b1 = s_meta.min_b[1], b2 = s_meta.min_b[2], m1 = s_meta.mul1, m2 = s_meta.mul2;
        const bool ovf = s_meta.overflow != 0;
        const unsigned max_key = ovf ? (unsigned)(n - 1) : (unsigned)((long long)s_meta.div_b[0] * s_meta.div_b[1] * s_meta.div_b[2] - 1);
        bits = 32 - __clz(max_key | 1u);
        for (int i = tid; i < P; i += VGS_THREADS) {
            unsigned key = 0xffffffffu;
            if (i < n) {
                if (ovf) key = (unsigned)i;                              // every point its own voxel ⇒ output == input
                else {
                    const float4 p = pts[i];
                    const int i0 = (int)(floorf(p.x * inv) - (float)b0), i1 = (int)(floorf(p.y * inv) - (float)b1), i2 = (int)(floorf(p.z * inv) - (float)b2);
                    key = (unsigned)(i0 + i1 * m1 + i2 * m2);
                }
            }
            s_key[i] = key; s_idx[i] = (unsigned short)i;
        }
    }
    __syncthreads();
    // ---- 3. stable sort by key (ties keep ascending input index; the padding ties with the largest key and stays behind it) ----
    vgs_radix_sort<VGS_E>(s_key, s_idx, s_cnt, s_wsum, bits);
    // ---- 4. segment heads → exclusive scan (thread t owns the consecutive items [t * per, (t + 1) * per)) ----
    const int per = P / VGS_THREADS;
    const int i_begin = tid * per;
    unsigned local = 0;
    for (int u = 0; u < per; ++u) { const int i = i_begin + u; if (i < n && (i == 0 || s_key[i] != s_key[i - 1])) ++local; }
    unsigned incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(FULL, incl, o); if (lane_id() >= o) incl += v; }
    if (lane_id() == 31) s_wsum[warp_id()] = incl;
    __syncthreads();
    if (warp_id() == 0) {
        const unsigned v = lane_id() < VGS_THREADS / 32 ? s_wsum[lane_id()] : 0u; unsigned inc2 = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_up_sync(FULL, inc2, o); if (lane_id() >= o) inc2 += u; }
        if (lane_id() < VGS_THREADS / 32) s_wsum[lane_id()] = inc2 - v;  // exclusive warp offsets
        if (lane_id() == 31) s_total = inc2;
    }
    __syncthreads();
    unsigned pos = s_wsum[warp_id()] + incl - local;
    for (int u = 0; u < per; ++u) { const int i = i_begin + u; if (i < n && (i == 0 || s_key[i] != s_key[i - 1])) seg_start[pos++] = (unsigned)i; }
    const int nseg = (int)s_total;
    if (tid == 0) { seg_start[nseg] = (unsigned)n; *n_out_dev = nseg; }
    __syncthreads();
    // ---- 5. centroids: one thread per voxel, sequential fp32 sums in ascending input index, divided by (float)count ----
    for (int s = tid; s < nseg; s += VGS_THREADS) {
        const unsigned b = seg_start[s], e = seg_start[s + 1];
        float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
        for (unsigned k = b; k < e; ++k) {
            const unsigned idx = s_idx[k];
            const float4 p = __ldg(pts + idx);
            sx += p.x; sy += p.y; sz += p.z; si += p.w;
            if (membership) membership[idx] = s;
        }
        const float c = (float)(e - b);
        out[s] = make_float4(sx / c, sy / c, sz / c, si / c);
        if (out_keys) out_keys[s] = (int)s_key[b];
    }
}

// out must hold cnt.bound points.  n_out_dev receives the voxel count (device int).
inline int voxel_grid_device(const float4* in, Count cnt, float leaf, float4* out, int* n_out_dev, int* membership, int* out_keys,
                             VoxelGridWork& w, cudaStream_t s) {
    const int nb = cnt.bound;
    if (nb <= 0) { CUDA_TRY(cudaMemsetAsync(n_out_dev, 0, sizeof(int), s)); return LIORF_OK; }
    int rc;
    if (nb <= VGS_CAP && !w.force_large) {
        if ((rc = w.seg_start.reserve((size_t)nb + 1))) return rc;
        if (!w.small_attr_set) { CUDA_TRY(cudaFuncSetAttribute(k_vg_small, cudaFuncAttributeMaxDynamicSharedMemorySize, VGS_SMEM)); w.small_attr_set = true; }
        k_vg_small<<<1, VGS_THREADS, VGS_SMEM, s>>>(in, cnt, leaf, out, n_out_dev, membership, out_keys, w.meta, w.seg_start.p);
        CUDA_TRY(cudaGetLastError());
        return LIORF_OK;
    }
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

// localmap.cuh — GPU-resident local map: keyframe transform+concat (extractCloud, src/mapOptmization.cpp:1012-1044,
// transformPointCloud :310-329) and the uniform voxel-hash grid that replaces the per-frame FLANN kd-tree build
// (kdtreeSurfFromMap->setInputCloud, :1302).
//
// Grid: cell edge is exactly 1.0 m = sqrt of the reference's hard-coded acceptance radius (pointSearchSqDis[4] < 1.0,
// :1097), so every map point that can appear in an ACCEPTED 5-NN set lies in the 3x3x3 cells around the query cell.
// Cell coordinates are hashed by wrapping into a power-of-two torus (DX x DY x DZ); aliased cells only add far
// candidates, so the search stays exact.  Cells are stored x-fastest so the three x-neighbours of a row are one
// contiguous range of the cell-sorted point array.
#pragma once
#include "prims.cuh"

namespace liorf {

struct GridDims { int DX, DY, DZ; };       // powers of two
__host__ __device__ __forceinline__ int grid_cells(GridDims g) { return g.DX * g.DY * g.DZ; }
__device__ __forceinline__ int grid_cell_of(GridDims g, float x, float y, float z) {
    int cx = (int)floorf(x), cy = (int)floorf(y), cz = (int)floorf(z);
    return ((cz & (g.DZ - 1)) * g.DY + (cy & (g.DY - 1))) * g.DX + (cx & (g.DX - 1));
}

struct MapGrid {
    GridDims dims{256, 256, 32};
    DevBuf<unsigned> counts;       // ncells, all zero between builds
    DevBuf<unsigned> cell_start;   // ncells + 1 (exclusive prefix + sentinel)
    DevBuf<float4> sorted;         // cell-sorted copy: x,y,z, w = bit pattern of the ORIGINAL map index
    ScanWork scan;
    bool counts_clean = false;
};

__global__ void __launch_bounds__(256) k_grid_count(const float4* __restrict__ map, Count cnt, GridDims g, unsigned* __restrict__ counts) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cnt.get()) return;
    float4 p = map[i];
    atomicAdd(&counts[grid_cell_of(g, p.x, p.y, p.z)], 1u);
}
struct GridScanLoad { const unsigned* counts; __device__ __forceinline__ unsigned operator()(int i) const { return counts[i]; } };
struct GridScanStore {
    unsigned* cell_start; int ncells;
    __device__ __forceinline__ void operator()(int i, unsigned v, unsigned excl) const {
        cell_start[i] = excl;
        if (i == ncells - 1) cell_start[ncells] = excl + v;
    }
};
// Exclusive scan of the cell counts → cell_start, specialised for this array (2 M cells for the default 256 x 256 x 32 torus: 8 MB in,
// 8 MB out, most of it zeros).  Same single-pass decoupled look-back as k_scan_lookback (ticket + epoch-tagged status words, so the
// kernel replays from a CUDA graph), but a thread moves 16-byte vectors at consecutive addresses: a tile is four sub-tiles of
// 512 x 4 cells, each loaded / stored with perfectly coalesced uint4 accesses (the generic kernel reads 16 scalars per thread at a
// 64-byte stride: 25 us for this scan, most of the grid build).
constexpr int GSCAN_BLOCK = 512;
constexpr int GSCAN_SUB = 4;
constexpr int GSCAN_TILE = GSCAN_BLOCK * 4 * GSCAN_SUB;        // 8192 cells
__global__ void __launch_bounds__(GSCAN_BLOCK) k_grid_scan(const unsigned* __restrict__ counts, unsigned* __restrict__ cell_start, int nc,
                                                           unsigned long long* ticket, unsigned long long* status, int* err_flag) {
    __shared__ int s_tile;
    __shared__ unsigned s_epoch;
    __shared__ unsigned s_wsum[GSCAN_SUB][GSCAN_BLOCK / 32];
    __shared__ unsigned s_sub[GSCAN_SUB + 1];
    __shared__ unsigned s_prefix;
    if (threadIdx.x == 0) { unsigned e; s_tile = draw_ticket(ticket, e); s_epoch = e; }
    __syncthreads();
    const int tile = s_tile; const unsigned epoch = s_epoch;
    const int ntiles = (nc + GSCAN_TILE - 1) / GSCAN_TILE;
    if (tile >= ntiles) return;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    uint4 v[GSCAN_SUB]; unsigned incl[GSCAN_SUB], tsum[GSCAN_SUB];
#pragma unroll
    for (int k = 0; k < GSCAN_SUB; ++k) {
        const int i = tile * GSCAN_TILE + k * (GSCAN_BLOCK * 4) + threadIdx.x * 4;
        if (i + 3 < nc) v[k] = *reinterpret_cast<const uint4*>(counts + i);
        else { v[k].x = i < nc ? counts[i] : 0u; v[k].y = i + 1 < nc ? counts[i + 1] : 0u; v[k].z = i + 2 < nc ? counts[i + 2] : 0u; v[k].w = 0u; }
        tsum[k] = v[k].x + v[k].y + v[k].z + v[k].w;
        incl[k] = tsum[k];
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
        for (int k = 0; k < GSCAN_SUB; ++k) { const unsigned t = __shfl_up_sync(FULL, incl[k], o); if (l >= o) incl[k] += t; }
    }
    if (l == 31) {
#pragma unroll
        for (int k = 0; k < GSCAN_SUB; ++k) s_wsum[k][w] = incl[k];
    }
    __syncthreads();
    if (w < GSCAN_SUB) {                                  // warp k scans the 16 warp totals of sub-tile k
        const unsigned t = l < GSCAN_BLOCK / 32 ? s_wsum[w][l] : 0u;
        unsigned ti = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_up_sync(FULL, ti, o); if (l >= o) ti += u; }
        if (l < GSCAN_BLOCK / 32) s_wsum[w][l] = ti - t;
        if (l == GSCAN_BLOCK / 32 - 1) s_sub[w + 1] = ti;  // total of sub-tile w
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        unsigned run = 0, total = 0;
        if (threadIdx.x == 0) { s_sub[0] = 0; for (int k = 1; k <= GSCAN_SUB; ++k) { run += s_sub[k]; s_sub[k] = run; } }   // exclusive bases of the sub-tiles, [GSCAN_SUB] = tile total
        __syncwarp();
        total = s_sub[GSCAN_SUB];
        const unsigned pf = lookback_warp(status, tile, total, epoch, err_flag);
        if (threadIdx.x == 0) s_prefix = pf;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GSCAN_SUB; ++k) {
        const int i = tile * GSCAN_TILE + k * (GSCAN_BLOCK * 4) + threadIdx.x * 4;
        const unsigned e0 = s_prefix + s_sub[k] + s_wsum[k][w] + incl[k] - tsum[k];
        uint4 o; o.x = e0; o.y = e0 + v[k].x; o.z = o.y + v[k].y; o.w = o.z + v[k].z;
        if (i + 3 < nc) *reinterpret_cast<uint4*>(cell_start + i) = o;
        else { if (i < nc) cell_start[i] = o.x; if (i + 1 < nc) cell_start[i + 1] = o.y; if (i + 2 < nc) cell_start[i + 2] = o.z; }
        if (i <= nc - 1 && nc - 1 <= i + 3) {              // the thread that holds the last cell also writes the sentinel
            const int j = nc - 1 - i;
            const unsigned last_excl = j == 0 ? o.x : j == 1 ? o.y : j == 2 ? o.z : o.w;
            const unsigned last_v = j == 0 ? v[k].x : j == 1 ? v[k].y : j == 2 ? v[k].z : v[k].w;
            cell_start[nc] = last_excl + last_v;
        }
    }
}

// fills each cell from its end; counts return to zero, ready for the next build
__global__ void __launch_bounds__(256) k_grid_scatter(const float4* __restrict__ map, Count cnt, GridDims g, unsigned* __restrict__ counts,
                                                     const unsigned* __restrict__ cell_start, float4* __restrict__ sorted) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cnt.get()) return;
    float4 p = map[i];
    int c = grid_cell_of(g, p.x, p.y, p.z);
    unsigned k = atomicSub(&counts[c], 1u) - 1u;
    sorted[cell_start[c] + k] = make_float4(p.x, p.y, p.z, __int_as_float(i));
}

inline int build_map_grid(const float4* map, Count cnt, MapGrid& G, cudaStream_t s) {
    const int nc = grid_cells(G.dims);
    int rc;
    if ((rc = G.counts.reserve(nc))) return rc;
    if ((rc = G.cell_start.reserve((size_t)nc + 1))) return rc;
    if ((rc = G.sorted.reserve(cnt.bound > 0 ? cnt.bound : 1))) return rc;
    if (!G.counts_clean) { CUDA_TRY(cudaMemsetAsync(G.counts.p, 0, (size_t)nc * sizeof(unsigned), s)); G.counts_clean = true; }
    if (cnt.bound > 0) k_grid_count<<<(cnt.bound + 255) / 256, 256, 0, s>>>(map, cnt, G.dims, G.counts.p);
    {
        const int ntiles = (nc + GSCAN_TILE - 1) / GSCAN_TILE;
        if ((rc = reserve_zeroed(G.scan.status, ntiles, s))) return rc;
        k_grid_scan<<<ntiles, GSCAN_BLOCK, 0, s>>>(G.counts.p, G.cell_start.p, nc, G.scan.ticket, G.scan.status.p, G.scan.err_flag);
    }
    if (cnt.bound > 0) k_grid_scatter<<<(cnt.bound + 255) / 256, 256, 0, s>>>(map, cnt, G.dims, G.counts.p, G.cell_start.p, G.sorted.p);
    CUDA_TRY(cudaGetLastError());
    return LIORF_OK;
}

// ---- extractCloud: transform each selected keyframe by its pose and concatenate in selection order ----
struct KfSel { int src_off, count, dst_off, pad; float t[12]; };       // 64 B

__device__ __forceinline__ void transform_concat_body(const float4* __restrict__ kf_points, const KfSel* __restrict__ sel, int nsel,
                                                      int total, float4* __restrict__ out) {
    __shared__ int s_dst[1024];
    const int ns = nsel < 1024 ? nsel : 1024;
    for (int i = threadIdx.x; i < ns; i += blockDim.x) s_dst[i] = sel[i].dst_off;
    __syncthreads();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int lo = 0, hi = nsel - 1;                       // last selection with dst_off <= i
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        int d = mid < 1024 ? s_dst[mid] : sel[mid].dst_off;
        if (d <= i) lo = mid; else hi = mid - 1;
    }
    const KfSel& k = sel[lo];
    float4 p = kf_points[k.src_off + (i - k.dst_off)];
    out[i] = apply_affine_dev(k.t, p);
}
__global__ void __launch_bounds__(256) k_transform_concat(const float4* __restrict__ kf_points, const KfSel* __restrict__ sel, int nsel,
                                                          int total, float4* __restrict__ out) {
    transform_concat_body(kf_points, sel, nsel, total, out);
}
// Same, with the selection count and the point total read from a header entry in front of the table (hdr[0].src_off = nsel,
// hdr[0].count = total, selections from hdr[1]): no per-call kernel argument, so the local-map chain replays from a CUDA graph.
__global__ void __launch_bounds__(256) k_transform_concat_hdr(const float4* __restrict__ kf_points, const KfSel* __restrict__ hdr, float4* __restrict__ out) {
    transform_concat_body(kf_points, hdr + 1, hdr[0].src_off, hdr[0].count, out);
}

}  // namespace liorf

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
    if ((rc = launch_scan<512, 16>(Count::of_host(nc), GridScanLoad{G.counts.p}, GridScanStore{G.cell_start.p, nc}, G.scan, nullptr, s))) return rc;
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

// sc_shard.cuh — ScanContext search over a keyframe database sharded across GPUs, exchanged through NVLink PEER MEMORY.
// (SURVEY §8e / BASELINE config 5: rows [g K/G, (g+1) K/G) of the database on rank g, queries replicated.)
//
// Every rank owns an exchange WINDOW in its HBM that all peers map (cudaIpc across processes, plain pointers inside one).
// A phase's producer kernel PUSHES its small per-query result into slot [my rank] of every rank's window with ordinary
// stores over NVLink, fences at system scope and raises flag[my rank][phase] = batch number in every window; the consumer
// kernel of that phase spins on its OWN window's flags (local L2 reads) before it touches the slots.  No NCCL call, no host
// round trip, no collective launch: a batch is one stream of kernels per rank, and the transfer of a phase overlaps whatever
// the peers are still computing.
//
//   phase K: ring keys of this rank's SLICE of the queries (every rank would otherwise read all Q descriptors, 9.6 kB each, to derive them)
//   phase T: the three smallest tile minima of the tensor-core filter per query, each inflated by this rank's error bound, so
//            that every rank can derive the GLOBAL candidate threshold (a rank that only knew its local third-smallest value
//            would re-rank ~100 candidates per query whatever the shard size; with the global bound the re-rank work shards too)
//   phase C: exact local top-3 {f32 dist, i32 global idx} → every rank merges the G lists by (dist, idx): identical global top-3
//   phase D: distanceBtnScanContext of the candidates THIS rank owns {f64 dist, i32 shift}, written by the stage-2 kernel itself into every
//            window (compute and transfer in ONE kernel); the consumer reads the owner's slot of each pair → decision
// Exactly the global top-3 are evaluated, so results equal the unsharded search bit for bit (tests/test_gpu_sc_shard.py).
//
// Reuse without double buffering is safe: a rank enters phase p of batch b+1 only after it has consumed phase D of batch b,
// which needs every rank's phase-D push, which each rank issues after it has finished reading phases T and C of batch b.
#pragma once
#include "sc_tensor.cuh"

namespace liorf {

constexpr int SCSH_MAX = 16;                 // ranks
enum { SCSH_T = 0, SCSH_C = 1, SCSH_D = 2, SCSH_K = 3 };

struct ShardWin {                            // passed by value to the kernels
    unsigned char* base[SCSH_MAX];           // window of every rank as mapped into THIS process (base[rank] = own)
    int rank, world;
    int row_begin[SCSH_MAX + 1];             // database rows [row_begin[g], row_begin[g + 1]) live on rank g
    unsigned long long* wait_ns;             // optional [4]: nanoseconds block 0 spent waiting for the peers, per phase (accumulated; %globaltimer)
    unsigned long long off[4], stride[4];    // byte offset of a phase's region inside a window, byte stride between source slots (phase K: one shared array, stride 0)
};
__host__ __device__ __forceinline__ size_t scsh_flag_off(int src, int phase) { return ((size_t)src * 4 + phase) * sizeof(unsigned); }

// consumer side: thread 0 waits until every peer's flag of `phase` has reached `batch` (own window, acquire at system scope)
__device__ __forceinline__ void scsh_wait(const ShardWin& W, int phase, const unsigned* batch_p, int* err_flag) {
    const unsigned batch = *batch_p;
    if (threadIdx.x == 0) {
        unsigned long long t0 = 0;
        if (W.wait_ns && blockIdx.x == 0 && blockIdx.y == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
        for (int g = 0; g < W.world; ++g) {
            if (g == W.rank) continue;
            const unsigned* f = reinterpret_cast<const unsigned*>(W.base[W.rank] + scsh_flag_off(g, phase));
            unsigned v, spins = 0;
            while (true) {
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
                if ((int)(v - batch) >= 0) break;
                // a peer that never arrives is an error, not a hang: give up after ~10 s, and at once if the error flag is already up (a dead
                // peer must not cost this bound again in every later wait)
                if ((++spins & 1023u) == 0u && (spins > (1u << 24) || *reinterpret_cast<volatile int*>(err_flag) != 0)) { atomicExch(err_flag, 3); break; }
            }
        }
        if (t0) { unsigned long long t1; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1)); W.wait_ns[phase] += t1 - t0; }
    }
    __syncthreads();
}

// the batch number lives in device memory (bumped by the first kernel of a batch) so that a batch is the same kernel sequence with the
// same arguments every time — it replays from a CUDA graph
__global__ void k_scsh_next_batch(unsigned* batch) { *batch += 1u; }

// producer side: copy `words` 32-bit words from src into slot [my rank] of every window, then (last block) raise the flags
__global__ void __launch_bounds__(256) k_scsh_push(ShardWin W, int phase, const unsigned* __restrict__ src, size_t words, const unsigned* __restrict__ batch_p, unsigned* counter,
                                                  size_t dst_byte_off = 0) {
    const size_t slot = W.off[phase] + (size_t)W.rank * W.stride[phase] + dst_byte_off;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned v = src[i];
        for (int g = 0; g < W.world; ++g) reinterpret_cast<unsigned*>(W.base[g] + slot)[i] = v;
    }
    __threadfence_system();
    __syncthreads();
    __shared__ bool s_last;
    if (threadIdx.x == 0) { const unsigned t = atomicAdd(counter, 1u); s_last = (t == gridDim.x - 1); }
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x == 0) *counter = 0u;
    __threadfence_system();
    if ((int)threadIdx.x < W.world && (int)threadIdx.x != W.rank) {
        unsigned* f = reinterpret_cast<unsigned*>(W.base[threadIdx.x] + scsh_flag_off(W.rank, phase));
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(*batch_p) : "memory");
    }
}

// phase K consumer: the ring keys of all queries, derived slice by slice on the ranks, are complete in this window → contiguous copy
__global__ void __launch_bounds__(256) k_scsh_gather_keys(ShardWin W, const unsigned* __restrict__ batch, int words, unsigned* __restrict__ dst, int* err_flag) {
    scsh_wait(W, SCSH_K, batch, err_flag);
    const unsigned* src = reinterpret_cast<const unsigned*>(W.base[W.rank] + W.off[SCSH_K]);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < words; i += gridDim.x * blockDim.x) dst[i] = __ldcg(src + i);
}

// phase T producer: merge the per-split partial top-3 tile minima of a query, inflate them by this rank's error bound —
// u_j = m_j + eps_r(q) is an upper bound of the TRUE distance of a distinct database key — and store them straight into slot
// [my rank] of every window; the last block raises the flags (compute + transfer in one kernel)
__global__ void __launch_bounds__(128) k_scsh_u3_push(ShardWin W, const float* __restrict__ part, int n_rows, const float* __restrict__ qnorm, int Q,
                                                     const unsigned* __restrict__ nmax_bits, const unsigned* __restrict__ batch_p, unsigned* counter) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < Q) {
        float t1 = 3.0e38f, t2 = 3.0e38f, t3 = 3.0e38f;
#pragma unroll
        for (int sp = 0; sp < SCS_SPLITS; ++sp) {
            const float* p = part + ((size_t)sp * n_rows + q) * 3;
            top3_min_update(p[0], t1, t2, t3); top3_min_update(p[1], t1, t2, t3); top3_min_update(p[2], t1, t2, t3);
        }
        const float eps = SCT_EPS_REL * (qnorm[q] + __uint_as_float(*nmax_bits));
        const float u1 = t1 < 1.0e38f ? t1 + eps : 3.0e38f, u2 = t2 < 1.0e38f ? t2 + eps : 3.0e38f, u3 = t3 < 1.0e38f ? t3 + eps : 3.0e38f;
        for (int g = 0; g < W.world; ++g) {
            float* o = reinterpret_cast<float*>(W.base[g] + W.off[SCSH_T] + (size_t)W.rank * W.stride[SCSH_T]) + 3 * (size_t)q;
            o[0] = u1; o[1] = u2; o[2] = u3;
        }
    }
    __threadfence_system();
    __syncthreads();
    __shared__ bool s_last;
    if (threadIdx.x == 0) { const unsigned t = atomicAdd(counter, 1u); s_last = (t == gridDim.x - 1); }
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x == 0) *counter = 0u;
    __threadfence_system();
    if ((int)threadIdx.x < W.world && (int)threadIdx.x != W.rank) {
        unsigned* f = reinterpret_cast<unsigned*>(W.base[threadIdx.x] + scsh_flag_off(W.rank, SCSH_T));
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(*batch_p) : "memory");
    }
}
// phase T consumer: U3(q) = third smallest over all ranks' bounds >= the true global third-smallest distance d3; a key of THIS
// rank in the global top-3 has d~ <= d + eps_r <= d3 + eps_r <= U3 + eps_r =: thr(q)
__global__ void __launch_bounds__(128) k_scsh_thr(ShardWin W, const unsigned* __restrict__ batch, const float* __restrict__ qnorm, int Q, const unsigned* __restrict__ nmax_bits,
                                                 float* __restrict__ thr, int* err_flag) {
    scsh_wait(W, SCSH_T, batch, err_flag);
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    float t1 = 3.0e38f, t2 = 3.0e38f, t3 = 3.0e38f;
    for (int g = 0; g < W.world; ++g) {
        const float* p = reinterpret_cast<const float*>(W.base[W.rank] + W.off[SCSH_T] + (size_t)g * W.stride[SCSH_T]) + 3 * (size_t)q;
        top3_min_update(__ldcg(p), t1, t2, t3); top3_min_update(__ldcg(p + 1), t1, t2, t3); top3_min_update(__ldcg(p + 2), t1, t2, t3);
    }
    const float eps = SCT_EPS_REL * (qnorm[q] + __uint_as_float(*nmax_bits));
    thr[q] = t3 < 1.0e38f ? t3 + eps : 3.0e38f;
}

// phase C consumer: the G local top-3 lists → global top-3 by (dist, idx)   [same arithmetic as k_sc_merge_top3]; the pairs whose
// candidate THIS rank owns go to a compact list on the way (order irrelevant: stage-2 results are keyed by the pair id)
__global__ void __launch_bounds__(128) k_scsh_merge(ShardWin W, const unsigned* __restrict__ batch, int Q, float* __restrict__ out_d, int* __restrict__ out_i, int* err_flag,
                                                   int own_begin, int own_count, int* __restrict__ list, int* __restrict__ n_list) {
    scsh_wait(W, SCSH_C, batch, err_flag);
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    Top3 t; top3_init(t);
    if (q < Q) {
        for (int g = 0; g < W.world; ++g) {
            const float* pd = reinterpret_cast<const float*>(W.base[W.rank] + W.off[SCSH_C] + (size_t)g * W.stride[SCSH_C]);
            const int* pi = reinterpret_cast<const int*>(pd + 3 * (size_t)Q);
#pragma unroll
            for (int j = 0; j < 3; ++j) { const int i = __ldcg(pi + 3 * (size_t)q + j); if (i != 0x7fffffff) top3_insert(t, __ldcg(pd + 3 * (size_t)q + j), i); }
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) { out_d[3 * (size_t)q + j] = t.d[j]; out_i[3 * (size_t)q + j] = t.i[j]; }
    }
    if (!list) return;
    int n_own = 0; bool own[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) { own[j] = q < Q && t.i[j] != 0x7fffffff && t.i[j] - own_begin >= 0 && t.i[j] - own_begin < own_count; n_own += own[j] ? 1 : 0; }
    int incl = n_own;                                        // warp-aggregated append: one atomic per warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, incl, o); if (lane_id() >= o) incl += v; }
    const int total = __shfl_sync(FULL, incl, 31);
    int base = 0;
    if (lane_id() == 31 && total > 0) base = atomicAdd(n_list, total);
    base = __shfl_sync(FULL, base, 31) + incl - n_own;
#pragma unroll
    for (int j = 0; j < 3; ++j) if (own[j]) list[base++] = 3 * q + j;
}

// sector key and column norms (k_sc_keys_batch arithmetic) of the queries of the pairs THIS rank owns, one CTA per list entry (a query
// with two owned candidates is simply written twice with the same values); the other queries' descriptors are never read here
__global__ void __launch_bounds__(64) k_scsh_skcn_owned(const double* __restrict__ desc, const int* __restrict__ list, const int* __restrict__ n_list,
                                                        double* __restrict__ sectorkey, double* __restrict__ colnorm) {
    const int n = *n_list;
    const int t = threadIdx.x;
    for (int it = blockIdx.x; it < n; it += gridDim.x) {
        const int e = list[it] / 3;
        if (t < SC_SECTOR) {
            double s = 0, q = 0;
            for (int r = 0; r < SC_RING; ++r) { const double v = desc[(size_t)e * SC_DESC + r * SC_SECTOR + t]; s += v; q += v * v; }
            sectorkey[(size_t)e * SC_SECTOR + t] = s / SC_RING;
            colnorm[(size_t)e * SC_SECTOR + t] = sqrt(q);
        }
    }
}

// what the stage-2 kernel needs to push its results itself (sc_distance.cuh): every owned pair's {f64 dist, i32 shift} goes straight
// from the warp that computed it into slot [my rank] of EVERY window; the last block raises the phase-D flags
struct ShardPush { int enabled; int Q; ShardWin W; unsigned* counter; const unsigned* batch_p; };

// phase D consumer: each (query, candidate) pair was evaluated by the rank that owns the candidate's row → read THAT rank's slot;
// then the decision of detectLoopClosureID (:302-340): candidates in kNN order, strict <, threshold
__global__ void __launch_bounds__(128) k_scsh_decide(ShardWin W, const unsigned* __restrict__ batch, const int* __restrict__ cand, int Q, int* __restrict__ loop_id, int* __restrict__ shift,
                                                    double* __restrict__ dist, int* err_flag) {
    scsh_wait(W, SCSH_D, batch, err_flag);
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    double mn = 10000000.0; int al = 0, nn = 0;
#pragma unroll
    for (int c = 0; c < SC_NUM_CAND; ++c) {
        const size_t i = 3 * (size_t)q + c;
        const int idx = cand[i];
        double d = INFINITY; int sh = 0;
        if (idx != 0x7fffffff && idx >= 0) {
            int g = 0;
            while (g + 1 < W.world && idx >= W.row_begin[g + 1]) ++g;
            const unsigned char* slot = W.base[W.rank] + W.off[SCSH_D] + (size_t)g * W.stride[SCSH_D];
            d = __ldcg(reinterpret_cast<const double*>(slot) + i);
            sh = __ldcg(reinterpret_cast<const int*>(slot + (size_t)3 * Q * 8) + i);
        }
        if (d < mn) { mn = d; al = sh; nn = idx; }
    }
    loop_id[q] = mn < SC_DIST_THRES ? nn : -1; shift[q] = al; dist[q] = mn;
}

}  // namespace liorf

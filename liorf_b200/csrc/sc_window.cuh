// sc_window.cuh — the NVLink peer-memory exchange window of the sharded ScanContext search: layout, flag wait / raise primitives.
// Protocol and decomposition: see sc_shard.cuh.
#pragma once
#include "scancontext.cuh"

namespace liorf {

constexpr int SCSH_MAX = 16;                 // ranks
enum { SCSH_C = 0, SCSH_D = 1, SCSH_KEYS = 2, SCSH_NPHASE = 4 };

struct ShardWin {                            // passed by value to the kernels
    unsigned char* base[SCSH_MAX];           // window of every rank as mapped into THIS process (base[rank] = own)
    int rank, world;
    int nowait;                              // measurement only (tools/profile_sc_shard.py): consumers do not wait for the peers' flags
    int qmax;                                // capacity of the per-batch arrays (queries)
    int row_begin[SCSH_MAX + 1];             // database rows [row_begin[g], row_begin[g + 1]) live on rank g
    unsigned long long* wait_ns;             // optional [SCSH_NPHASE]: nanoseconds block 0 spent waiting for the peers, per phase (%globaltimer)
    unsigned long long off_c, off_d;         // byte offsets inside a window: C = { f32 dist[3 qmax]; i32 idx[3 qmax] }, D = { f64 dist[3 qmax]; i32 shift[3 qmax] }
    float* keys_all[SCSH_MAX];               // the replicated ring-key array [K_total][20] of every rank (own entry = local pointer)
};
__host__ __device__ __forceinline__ size_t scsh_flag_off(int src, int phase) { return ((size_t)src * SCSH_NPHASE + phase) * sizeof(unsigned); }
__device__ __forceinline__ float* scsh_c_dist(const ShardWin& W, int g) { return reinterpret_cast<float*>(W.base[g] + W.off_c); }
__device__ __forceinline__ int* scsh_c_idx(const ShardWin& W, int g) { return reinterpret_cast<int*>(W.base[g] + W.off_c) + (size_t)3 * W.qmax; }
__device__ __forceinline__ double* scsh_d_dist(const ShardWin& W, int g) { return reinterpret_cast<double*>(W.base[g] + W.off_d); }
__device__ __forceinline__ int* scsh_d_shift(const ShardWin& W, int g) { return reinterpret_cast<int*>(W.base[g] + W.off_d + (size_t)3 * W.qmax * 8); }

// consumer side: warp 0 waits until every peer's flag of `phase` has reached `target` — one lane per peer, all flags polled in
// parallel in this rank's OWN window (local L2 reads, acquire at system scope)
__device__ __forceinline__ void scsh_wait(const ShardWin& W, int phase, unsigned target, int* err_flag) {
    if (!W.nowait && threadIdx.x < 32) {
        unsigned long long t0 = 0;
        const bool timed = W.wait_ns && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0;
        if (timed) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
        const int g = threadIdx.x;
        if (g < W.world && g != W.rank) {
            const unsigned* f = reinterpret_cast<const unsigned*>(W.base[W.rank] + scsh_flag_off(g, phase));
            unsigned v, spins = 0;
            while (true) {
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
                if ((int)(v - target) >= 0) break;
                // a peer that never arrives is an error, not a hang: give up after ~10 s, and at once if the error flag is already up (a dead
                // peer must not cost this bound again in every later wait)
                if ((++spins & 1023u) == 0u && (spins > (1u << 25) || *reinterpret_cast<volatile int*>(err_flag) != 0)) { atomicCAS(err_flag, 0, 0x30 + phase); break; }   // 0x30 + phase: a peer's flag never came
            }
        }
        __syncwarp();
        if (timed) { unsigned long long t1; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1)); W.wait_ns[phase] += t1 - t0; }
    }
    __syncthreads();
}

// producer side, end of a pushing kernel: every block has fenced its stores at system scope; the LAST block to get here raises
// flag[my rank][phase] = value in every peer's window (one lane per peer).  Must be reached by all threads of every block.
__device__ __forceinline__ void scsh_raise(const ShardWin& W, int phase, unsigned value, unsigned* counter) {
    __threadfence_system();
    __syncthreads();
    __shared__ bool s_scsh_last;
    if (threadIdx.x == 0) { const unsigned t = atomicAdd(counter, 1u); s_scsh_last = (t == gridDim.x * gridDim.y - 1); }
    __syncthreads();
    if (!s_scsh_last) return;
    if (threadIdx.x == 0) *counter = 0u;
    __threadfence_system();
    if ((int)threadIdx.x < W.world && (int)threadIdx.x != W.rank) {
        unsigned* f = reinterpret_cast<unsigned*>(W.base[threadIdx.x] + scsh_flag_off(W.rank, phase));
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(value) : "memory");
    }
}

// what a pushing kernel needs (re-rank kernel: phase C; stage-2 kernel: phase D).  q0 = first query of this rank's slice.
struct ShardPush { int enabled; int q0; ShardWin W; unsigned* counter; const unsigned* batch_p; };

}  // namespace liorf

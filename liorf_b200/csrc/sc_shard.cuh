// sc_shard.cuh — ScanContext search over a keyframe database sharded across GPUs, exchanged through NVLink PEER MEMORY.
// (SURVEY §8e / BASELINE config 5.)
//
// Decomposition ("replicated index, sharded payload" — the reference keeps the same two structures apart: the ring-key
// kd-tree `polarcontext_invkeys_mat_` / `polarcontext_tree_` is a search index over the descriptors `polarcontexts_`,
// include/Scancontext.h:104-113):
//   * the PAYLOAD of database row k — descriptor (9 600 B), sector key (480 B), column norms (480 B) — lives only on the
//     rank that owns the row: rows [row_begin[g], row_begin[g+1]) on rank g;
//   * the INDEX — the 80-byte fp32 ring key of every row (0.75 % of a row) — is replicated: each rank pushes the keys of its rows
//     into every rank's key array over NVLink once, when the database was loaded or has grown (k_scsh_push_keys);
//   * stage 1 (ring-key top-3, detectLoopClosureID :289-295) is sharded by QUERY: rank g answers queries
//     [g Q/G, (g+1) Q/G) against ALL keys, so every O(Q) step of the stage (ring keys of the queries, operand images, GEMM
//     filter, thresholds, exact re-rank) shrinks with the rank count and needs no exchange at all — the result IS the global top-3;
//   * stage 2 (distanceBtnScanContext, :302-317) is sharded by OWNER: the rank that holds a candidate's descriptor evaluates the pair.
// Two exchanges per batch, both done by the producing kernels themselves (plain stores into every rank's window, fence at
// system scope, release-store of flag[my rank][phase] = batch number; consumers spin on flags in their OWN window):
//   phase C: the exact global top-3 {f32 dist, i32 idx} of a rank's query slice, written by the re-rank kernel into rows
//            [3 q0, 3 q1) of the ONE candidate array of every window;
//   phase D: {f64 dist, i32 shift} of every pair a rank owns, written by the stage-2 warp that computed it into entry [pair] of
//            the ONE pair array of every window → every rank takes the decision of detectLoopClosureID (:319-340) for all queries.
// Exactly the global top-3 are evaluated, so loop ids / shifts / distances equal the unsharded search bit for bit.
//
// Reuse without double buffering is safe: a rank pushes phase C of batch b+1 only after its own decision kernel of batch b, which
// waited for every rank's phase-D flag, which each rank raises after its collect kernel copied phase C of batch b out of the
// window; and a rank pushes phase D of batch b+1 only after it has seen every rank's phase-C flag of batch b+1, which each rank
// raises after its decision kernel of batch b has read phase D.
#pragma once
#include "sc_tensor.cuh"

namespace liorf {

// ---- the replicated index: the ring keys of this rank's rows go into every rank's key array (once per database change) ----
__global__ void __launch_bounds__(256) k_scsh_push_keys(ShardWin W, const float* __restrict__ keys, int n_rows, unsigned gen, unsigned* counter) {
    const size_t words = (size_t)n_rows * SC_RING, dst0 = (size_t)W.row_begin[W.rank] * SC_RING;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (size_t)gridDim.x * blockDim.x) {
        const float v = keys[i];
        for (int g = 0; g < W.world; ++g) W.keys_all[g][dst0 + i] = v;
    }
    scsh_raise(W, SCSH_KEYS, gen, counter);
}
__global__ void k_scsh_wait_keys(ShardWin W, unsigned gen, int* err_flag) { scsh_wait(W, SCSH_KEYS, gen, err_flag); }
// The consumer side of a phase is ONE WARP: a <<<1, 32>>> kernel in front of the kernel that reads the exchanged data.  Waiting inside the
// consumer kernels themselves (every CTA of k_scsh_collect / k_scsh_decide spinning) was how round 1 and the first version of this file did
// it — and with several query batches in flight per GPU the spinning CTAs of the lanes that wait (256 CTAs x 128 threads and their registers
// per lane at Q = 32 768) can leave no SM with room for the 320-thread / 213 kB tensor-core CTAs or the stage-2 CTAs of the lane whose push
// the OTHER GPUs are waiting for: a cross-GPU resource cycle, observed as soon as host copies skewed the lanes (bench --gpus 2 / 8, r2).
// One warp per waiting lane cannot starve anything.
__global__ void k_scsh_wait_phase(ShardWin W, int phase, const unsigned* __restrict__ batch, int* err_flag) { scsh_wait(W, phase, *batch, err_flag); }

// phase C producer of the CUDA-core search path (the tensor-core path pushes from its re-rank kernel): the exact global top-3 of this
// rank's query slice, n_q rows starting at query q0, into the candidate array of every window
__global__ void __launch_bounds__(256) k_scsh_push_c(ShardWin W, const float* __restrict__ d, const int* __restrict__ idx, int q0, int n_q, const unsigned* __restrict__ batch_p,
                                                    unsigned* counter) {
    const int words = 3 * n_q;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < words; i += gridDim.x * blockDim.x) {
        const float dv = d[i]; const int iv = idx[i];
        const size_t e = 3 * (size_t)q0 + i;
        for (int g = 0; g < W.world; ++g) { scsh_c_dist(W, g)[e] = dv; scsh_c_idx(W, g)[e] = iv; }
    }
    scsh_raise(W, SCSH_C, *batch_p, counter);
}

// phase C consumer (runs behind k_scsh_wait_phase(C)): every slice of the candidate array in this window is complete → local copy (the window is overwritten by the
// next batch) + the compact list of the pairs whose candidate THIS rank owns (order irrelevant: stage-2 results are keyed by the pair id)
__global__ void __launch_bounds__(128) k_scsh_collect(ShardWin W, const unsigned* __restrict__ batch, int Q, float* __restrict__ out_d, int* __restrict__ out_i, int* err_flag,
                                                     int own_begin, int own_count, int* __restrict__ list, int* __restrict__ n_list) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    int ci[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff};
    if (q < Q) {
        const float* pd = scsh_c_dist(W, W.rank) + 3 * (size_t)q;
        const int* pi = scsh_c_idx(W, W.rank) + 3 * (size_t)q;
#pragma unroll
        for (int j = 0; j < 3; ++j) { ci[j] = __ldcg(pi + j); out_d[3 * (size_t)q + j] = __ldcg(pd + j); out_i[3 * (size_t)q + j] = ci[j]; }
    }
    int n_own = 0; bool own[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) { own[j] = q < Q && ci[j] != 0x7fffffff && ci[j] - own_begin >= 0 && ci[j] - own_begin < own_count; n_own += own[j] ? 1 : 0; }
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

// phase D consumer (runs behind k_scsh_wait_phase(D)): entry [pair] of the pair array was written by the rank that owns the candidate's row; then the decision of
// detectLoopClosureID (:302-340): candidates in kNN order, strict <, threshold.  Re-arms the owned-pair counter for the next batch.
__global__ void __launch_bounds__(128) k_scsh_decide(ShardWin W, const unsigned* __restrict__ batch, const int* __restrict__ cand, int Q, int* __restrict__ loop_id, int* __restrict__ shift,
                                                    double* __restrict__ dist, int* err_flag, int* __restrict__ n_list) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q == 0) *n_list = 0;
    if (q >= Q) return;
    const double* pd = scsh_d_dist(W, W.rank); const int* ps = scsh_d_shift(W, W.rank);
    double mn = 10000000.0; int al = 0, nn = 0;
#pragma unroll
    for (int c = 0; c < SC_NUM_CAND; ++c) {
        const size_t i = 3 * (size_t)q + c;
        const int idx = cand[i];
        double d = INFINITY; int sh = 0;
        if (idx != 0x7fffffff && idx >= 0) { d = __ldcg(pd + i); sh = __ldcg(ps + i); }
        if (d < mn) { mn = d; al = sh; nn = idx; }
    }
    loop_id[q] = mn < SC_DIST_THRES ? nn : -1; shift[q] = al; dist[q] = mn;
}

}  // namespace liorf

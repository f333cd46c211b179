// sc_distance.cuh — stage 2 of detectLoopClosureID for query batches: distanceBtnScanContext (include/Scancontext.cpp:116-148)
// with the candidate descriptor staged by the TMA engine.
//
// A TEAM of two warps (64 threads, thread ↔ descriptor column / thread ↔ shift) per (query, candidate) pair, looping over pairs;
// four teams per CTA, four CTAs per SM (≤ 64 registers per thread) → 16 pairs and 32 warps in flight per SM.  The pair count in
// flight is bounded by shared memory (13.9 kB per pair), so giving each pair two warps doubles the warps the schedulers can pick from
// — the arithmetic is chains of dependent fp64 operations (the reference's sequential sums), which one warp per pair left
// 46 % of the issue slots unable to cover (ncu, round 1) — and halves the latency of a pair, which is what a rank of a sharded
// search with only ~12k owned pairs feels.
// The candidate's 20x60 fp64 descriptor (9 600 B, a random row of a database that lives in HBM) is fetched by ONE bulk asynchronous
// copy (cp.async.bulk → shared memory, completion on an mbarrier) while the team derives the query's sector key / column norms and
// runs fastAlignUsingVkey; the fine search then reads its 7 shifted column sets from shared memory.
// The arithmetic — every fp64 sum in the reference's sequential order — is the one of k_sc_distance (scancontext.cuh), which stays in
// use for the live single-query call; results are bit-identical (tests/test_gpu_sc_tensor.py, test_gpu_deskew_sc.py).
// Algorithmic traffic per pair: 9 600 B candidate + 9 600 B query descriptor (three pairs share a query: L2) + 960 B candidate keys/norms.
#pragma once
#include "sc_shard.cuh"

namespace liorf {

constexpr int SCDB_TEAMS = 4;                      // pairs in flight per CTA
constexpr int SCDB_TEAM_THREADS = 64;
constexpr int SCDB_THREADS = SCDB_TEAMS * SCDB_TEAM_THREADS;
constexpr int SCDB_TEAM_BYTES = SC_DESC * 8 + 2 * SC_SECTOR * 8 + 7 * SC_SECTOR * 8;      // candidate descriptor + two sector keys + 7 x 60 similarities = 13 920 B
constexpr int SCDB_SMEM = SCDB_TEAMS * SCDB_TEAM_BYTES + SCDB_TEAMS * 8 + SCDB_TEAMS * 32;

__device__ __forceinline__ void team_sync(int team) { asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(SCDB_TEAM_THREADS) : "memory"); }

__global__ void __launch_bounds__(SCDB_THREADS, 4) k_sc_distance_bulk(const double* __restrict__ qdesc, const int* __restrict__ cand, int n_pairs, int cand_per_query,
                                                                     const double* __restrict__ db_desc, const double* __restrict__ db_sk, const double* __restrict__ db_cn,
                                                                     int own_begin, int own_count, double* __restrict__ out_dist, int* __restrict__ out_shift, int* err_flag,
                                                                     const int* __restrict__ pair_list, const int* __restrict__ n_list, ShardPush P) {
    extern __shared__ __align__(16) unsigned char scdb_smem[];
    const int team = threadIdx.x / SCDB_TEAM_THREADS, t = threadIdx.x % SCDB_TEAM_THREADS, l = lane_id();
    double* s_sc2 = reinterpret_cast<double*>(scdb_smem + (size_t)team * SCDB_TEAM_BYTES);
    double* s_vk1 = s_sc2 + SC_DESC; double* s_vk2 = s_vk1 + SC_SECTOR;
    double (*s_sim)[SC_SECTOR] = reinterpret_cast<double (*)[SC_SECTOR]>(s_vk2 + SC_SECTOR);
    const uint32_t bar = smem_u32(scdb_smem + (size_t)SCDB_TEAMS * SCDB_TEAM_BYTES + 8 * team);
    // cross-warp slots of a team's argmin: {value, shift} per warp
    double* s_red = reinterpret_cast<double*>(scdb_smem + (size_t)SCDB_TEAMS * SCDB_TEAM_BYTES + 8 * SCDB_TEAMS + 32 * team);
    __shared__ int s_abort;
    if (threadIdx.x == 0) s_abort = 0;
    if (t == 0) mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint32_t parity = 0;
    const int n_teams = gridDim.x * SCDB_TEAMS;
    const int n_items = pair_list ? *n_list : n_pairs;               // sharded search: the compact list of the pairs this rank owns
    const bool col = t < SC_SECTOR;                                  // threads 60..63 of a team only take part in the barriers
    for (int it = blockIdx.x * SCDB_TEAMS + team; it < n_items; it += n_teams) {
        const int pair = pair_list ? pair_list[it] : it;
        const int q = pair / cand_per_query;
        const int c = cand[pair];
        if (c == 0x7fffffff || c < 0) { if (t == 0 && out_dist) { out_dist[pair] = INFINITY; out_shift[pair] = 0; } continue; }      // team-uniform
        const int lc = c - own_begin;
        if (lc < 0 || lc >= own_count) continue;
        const double* sc1 = qdesc + (size_t)q * SC_DESC;
        team_sync(team);                                                // every thread is done with the previous pair's buffers
        if (t == 0) { mbar_expect_tx(bar, SC_DESC * 8); bulk_g2s(smem_u32(s_sc2), db_desc + (size_t)lc * SC_DESC, SC_DESC * 8, bar); }
        // the candidate's sector key twice in a row (in the not yet used similarity buffer): circshift(vkey2, s)[k] = vk2d[k - s + 60]
        // is then a plain offset from a per-thread base — no modular index arithmetic inside the 60-step sums
        double* s_vk2d = &s_sim[0][0];
        // the QUERY's sector key and column norms are derived here, from the descriptor the fine search reads anyway (thread ↔ column,
        // 20 independent loads; k_sc_keys_batch's arithmetic: sequential sums over the rings, :214-227, :75-81) — no per-query key
        // arrays, no kernel in front of this one
        double n1 = 0;
        if (col) {
            double sum = 0, sq = 0;
#pragma unroll
            for (int r = 0; r < SC_RING; ++r) { const double v = sc1[r * SC_SECTOR + t]; sum += v; sq += v * v; }
            s_vk1[t] = sum / SC_RING;
            n1 = sqrt(sq);
            const double v2 = db_sk[(size_t)lc * SC_SECTOR + t];
            s_vk2[t] = v2; s_vk2d[t] = v2; s_vk2d[t + SC_SECTOR] = v2;
        }
        team_sync(team);
        // fastAlignUsingVkey (:93-113): thread ↔ shift, sequential sum over columns; first strict minimum of the norm
        double best = 10000000.0; int best_s = 0x7fffffff;
        if (col) {
            const double* pA = s_vk2d + SC_SECTOR - t;
            double ss = 0;
#pragma unroll 12
            for (int k = 0; k < SC_SECTOR; ++k) { const double d = s_vk1[k] - pA[k]; ss += d * d; }
            const double nrm = sqrt(ss);
            if (nrm < best) { best = nrm; best_s = t; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ob = __shfl_xor_sync(FULL, best, o); int os = __shfl_xor_sync(FULL, best_s, o);
            if (ob < best || (ob == best && os < best_s)) { best = ob; best_s = os; }
        }
        if (l == 0) { s_red[2 * (t >> 5)] = best; reinterpret_cast<int*>(s_red + 2 * (t >> 5) + 1)[0] = best_s; }
        team_sync(team);                                                // also: every thread has finished reading the doubled sector key
        {
            const double b0 = s_red[0], b1 = s_red[2]; const int s0 = reinterpret_cast<const int*>(s_red + 1)[0], s1 = reinterpret_cast<const int*>(s_red + 3)[0];
            best = b0; best_s = s0;
            if (b1 < best || (b1 == best && s1 < best_s)) { best = b1; best_s = s1; }
        }
        const int align = best_s == 0x7fffffff ? 0 : best_s;            // every norm >= 1e7 or NaN ⇒ argmin stays 0 (:95)
        int shifts[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) shifts[j] = (align + j - SC_SEARCH_RADIUS + SC_SECTOR) % SC_SECTOR;
#pragma unroll
        for (int a = 1; a < 7; ++a) { int v = shifts[a]; int b = a - 1; while (b >= 0 && shifts[b] > v) { shifts[b + 1] = shifts[b]; --b; } shifts[b + 1] = v; }
        // the candidate descriptor has landed (or the wait gives up with the error flag set)
        if (!mbar_wait(bar, parity, &s_abort, err_flag)) break;        // error flag is set; the epilogue below still runs (flags, counter)
        parity ^= 1u;
        // fine search (:123-144): thread ↔ query column, dot over the 20 rings for each of the 7 shifts (7 independent chains, each in
        // ascending ring order), candidate columns from shared memory
        if (col) {
            int k2[7]; double dot[7];
#pragma unroll
            for (int j = 0; j < 7; ++j) { int v = t - shifts[j]; if (v < 0) v += SC_SECTOR; k2[j] = v; dot[j] = 0; }
#pragma unroll 5
            for (int r = 0; r < SC_RING; ++r) {
                const double a = sc1[r * SC_SECTOR + t];
#pragma unroll
                for (int j = 0; j < 7; ++j) dot[j] += a * s_sc2[r * SC_SECTOR + k2[j]];
            }
            const double* cn2 = db_cn + (size_t)lc * SC_SECTOR;
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                const double n2 = cn2[k2[j]];
                s_sim[j][t] = ((n1 == 0) | (n2 == 0)) ? -2.0 : dot[j] / (n1 * n2);      // distDirectSC :73-88
            }
        }
        team_sync(team);
        if (t < 32) {                                                    // first warp of the team: 7 sequential column sums, then the first strict minimum
            double dist = INFINITY; int sh = 0;
            if (l < 7) {
                double sum = 0; int eff = 0;
#pragma unroll 12
                for (int k = 0; k < SC_SECTOR; ++k) { double v = s_sim[l][k]; if (v != -2.0) { sum = sum + v; eff = eff + 1; } }     // the loads run ahead of the sequential adds
                dist = 1.0 - sum / eff;                                      // eff == 0 ⇒ NaN, never selected
                sh = shifts[0];
#pragma unroll
                for (int j = 1; j < 7; ++j) if (l == j) sh = shifts[j];
            }
            double mn = 10000000.0; int arg = 0;
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                double dj = __shfl_sync(FULL, dist, j); int sj = __shfl_sync(FULL, sh, j);
                if (dj < mn) { mn = dj; arg = sj; }
            }
            if (l == 0) {
                if (P.enabled) {             // push: entry [pair] of the pair array of every rank's window, straight from the warp that computed the pair
                    for (int g = 0; g < P.W.world; ++g) { scsh_d_dist(P.W, g)[pair] = mn; scsh_d_shift(P.W, g)[pair] = arg; }
                } else { out_dist[pair] = mn; out_shift[pair] = arg; }
            }
        }
    }
    if (P.enabled) scsh_raise(P.W, SCSH_D, *P.batch_p, P.counter);      // the last block to finish raises this rank's phase-D flag in every peer window
}

}  // namespace liorf

// sc_distance.cuh — stage 2 of detectLoopClosureID for query batches: distanceBtnScanContext (include/Scancontext.cpp:116-148)
// with the candidate descriptor staged by the TMA engine.
//
// One warp per (query, candidate) pair, looping over pairs.  The candidate's 20x60 fp64 descriptor (9 600 B, a random row
// of a database that lives in HBM) is fetched by ONE bulk asynchronous copy (cp.async.bulk → shared memory, completion on an
// mbarrier) while the warp runs fastAlignUsingVkey on the two sector keys; the fine search then reads the 7 shifted
// column sets from shared memory instead of issuing 140 dependent global loads per lane.  The arithmetic — every fp64 sum
// in the reference's sequential order — is the one of k_sc_distance (scancontext.cuh), which stays in use for the live
// single-query call; results are bit-identical (tests/test_gpu_sc_tensor.py, test_gpu_deskew_sc.py).
// Algorithmic traffic per pair: 9 600 B candidate + 9 600 B query descriptor (three pairs share a query: L2) + 960 B candidate keys/norms.
#pragma once
#include "sc_shard.cuh"

namespace liorf {

constexpr int SCDB_WARPS = 4;
constexpr int SCDB_WARP_BYTES = SC_DESC * 8 + 2 * SC_SECTOR * 8 + 7 * SC_SECTOR * 8;      // candidate descriptor + two sector keys + 7 x 60 similarities = 13 920 B
constexpr int SCDB_SMEM = SCDB_WARPS * SCDB_WARP_BYTES + SCDB_WARPS * 8;

__global__ void __launch_bounds__(SCDB_WARPS * 32) k_sc_distance_bulk(const double* __restrict__ qdesc, const int* __restrict__ cand, int n_pairs, int cand_per_query,
                                                                      const double* __restrict__ db_desc, const double* __restrict__ db_sk, const double* __restrict__ db_cn,
                                                                      int own_begin, int own_count, double* __restrict__ out_dist, int* __restrict__ out_shift, int* err_flag,
                                                                      const int* __restrict__ pair_list, const int* __restrict__ n_list, ShardPush P) {
    extern __shared__ __align__(16) unsigned char scdb_smem[];
    const int w = warp_id(), l = lane_id();
    double* s_sc2 = reinterpret_cast<double*>(scdb_smem + (size_t)w * SCDB_WARP_BYTES);
    double* s_vk1 = s_sc2 + SC_DESC; double* s_vk2 = s_vk1 + SC_SECTOR;
    double (*s_sim)[SC_SECTOR] = reinterpret_cast<double (*)[SC_SECTOR]>(s_vk2 + SC_SECTOR);
    const uint32_t bar = smem_u32(scdb_smem + (size_t)SCDB_WARPS * SCDB_WARP_BYTES + 8 * w);
    __shared__ int s_abort;
    if (threadIdx.x == 0) s_abort = 0;
    if (l == 0) mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint32_t parity = 0;
    const int n_warps = gridDim.x * SCDB_WARPS;
    const int n_items = pair_list ? *n_list : n_pairs;               // sharded search: the compact list of the pairs this rank owns
    for (int it = blockIdx.x * SCDB_WARPS + w; it < n_items; it += n_warps) {
        const int pair = pair_list ? pair_list[it] : it;
        const int q = pair / cand_per_query;
        const int c = cand[pair];
        if (c == 0x7fffffff || c < 0) { if (l == 0 && out_dist) { out_dist[pair] = INFINITY; out_shift[pair] = 0; } continue; }
        const int lc = c - own_begin;
        if (lc < 0 || lc >= own_count) continue;
        const double* sc1 = qdesc + (size_t)q * SC_DESC;
        const double* cn2 = db_cn + (size_t)lc * SC_SECTOR;
        __syncwarp();                                                   // every lane is done with the previous pair's buffers
        if (l == 0) { mbar_expect_tx(bar, SC_DESC * 8); bulk_g2s(smem_u32(s_sc2), db_desc + (size_t)lc * SC_DESC, SC_DESC * 8, bar); }
        // the candidate's sector key twice in a row (in the not yet used similarity buffer): circshift(vkey2, s)[k] = vk2d[k - s + 60]
        // is then a plain offset from a per-lane base — no modular index arithmetic inside the 60-step sums
        double* s_vk2d = &s_sim[0][0];
        // the QUERY's sector key and column norms are derived here, from the descriptor the fine search reads anyway (lane ↔ column,
        // the 40 loads of a lane are independent; k_sc_keys_batch's arithmetic: sequential sums over the rings, :214-227, :75-81) —
        // no per-query key arrays, no kernel in front of this one
        double n1A = 0, n1B = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = l + 32 * h;
            if (k < SC_SECTOR) {
                double sum = 0, sq = 0;
#pragma unroll
                for (int r = 0; r < SC_RING; ++r) { const double v = sc1[r * SC_SECTOR + k]; sum += v; sq += v * v; }
                s_vk1[k] = sum / SC_RING;
                if (h == 0) n1A = sqrt(sq); else n1B = sqrt(sq);
                const double v2 = db_sk[(size_t)lc * SC_SECTOR + k];
                s_vk2[k] = v2; s_vk2d[k] = v2; s_vk2d[k + SC_SECTOR] = v2;
            }
        }
        __syncwarp();
        // fastAlignUsingVkey (:93-113): lane ↔ shift, sequential sum over columns; first strict minimum of the norm
        double best = 10000000.0; int best_s = 0x7fffffff;
        {
            const int sA = l, sB = l + 32;
            const bool hasB = sB < SC_SECTOR;
            const double* pA = s_vk2d + SC_SECTOR - sA;
            const double* pB = s_vk2d + SC_SECTOR - (hasB ? sB : 0);
            double ssA = 0, ssB = 0;
#pragma unroll 12
            for (int k = 0; k < SC_SECTOR; ++k) {
                const double v1 = s_vk1[k];
                const double dA = v1 - pA[k], dB = v1 - pB[k];
                ssA += dA * dA; ssB += dB * dB;
            }
            const double nA = sqrt(ssA), nB = sqrt(ssB);
            if (nA < best) { best = nA; best_s = sA; }
            if (hasB && nB < best) { best = nB; best_s = sB; }
        }
        __syncwarp();                                                   // the similarity buffer is free again
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ob = __shfl_xor_sync(FULL, best, o); int os = __shfl_xor_sync(FULL, best_s, o);
            if (ob < best || (ob == best && os < best_s)) { best = ob; best_s = os; }
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
        // fine search (:123-144): lane ↔ column, dot over the 20 rings for each of the 7 shifts, candidate columns from shared memory
        for (int k = l; k < SC_SECTOR; k += 32) {
            double a1[SC_RING];
#pragma unroll
            for (int r = 0; r < SC_RING; ++r) a1[r] = sc1[r * SC_SECTOR + k];
            const double n1 = k < 32 ? n1A : n1B;
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                int k2 = k - shifts[j]; if (k2 < 0) k2 += SC_SECTOR;
                const double n2 = cn2[k2];
                double dot = 0;
#pragma unroll
                for (int r = 0; r < SC_RING; ++r) dot += a1[r] * s_sc2[r * SC_SECTOR + k2];
                s_sim[j][k] = ((n1 == 0) | (n2 == 0)) ? -2.0 : dot / (n1 * n2);      // distDirectSC :73-88
            }
        }
        __syncwarp();
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
    if (P.enabled) scsh_raise(P.W, SCSH_D, *P.batch_p, P.counter);      // the last block to finish raises this rank's phase-D flag in every peer window
}

}  // namespace liorf

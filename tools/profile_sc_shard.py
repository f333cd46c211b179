#!/usr/bin/env python
"""The sharded ScanContext search (csrc/sc_shard.cuh) on ONE GPU, for development without an 8-GPU box:
  1. G ranks emulated as contexts on one device (step-wise enqueue), one complete batch per lane → results checked against the
     unsharded search, every window filled with a complete set of phase-C / phase-D data;
  2. per-step CUDA-event times of rank 0 (its kernels have the GPU to themselves, as on its own device);
  3. throughput of RANK 0 ALONE (`liorf_sc_shard_debug_nowait`: consumers do not wait for flags, they read the complete data step 1
     left in the windows — the work of one rank of a G-rank search, minus the time spent waiting for peers) with `lanes` batches in
     flight, next to the unsharded search with the same number of lanes → predicted speed-up of G GPUs over one.
usage: python tools/profile_sc_shard.py [K] [Q] [world] [lanes] [reps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import liorf_b200
    from liorf_b200.sc_sharded import PeerShardedSearch
    from tools import synth
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    Q = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
    world = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    n_lanes = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    reps = int(sys.argv[5]) if len(sys.argv) > 5 else 10
    bounds = [K * g // world for g in range(world + 1)]
    dev = torch.device("cuda:0")

    def load(c, lo, hi):
        c.reserve(1024, 1024, 0, max(hi - lo, 1))
        for s in range(lo, hi, 10000):
            c.scAddDescriptors(synth.sc_descriptors(min(10000, hi - s), first=s))

    # lane sets: lanes[l][g] = (context, search) of rank g; lane 0 owns the shards, the others borrow them (and the replicated index)
    owners = [liorf_b200.Context() for _ in range(world)]
    for g, c in enumerate(owners):
        load(c, bounds[g], bounds[g + 1])
    lanes = [[(c, PeerShardedSearch(c, g, world, bounds, Q, torch)) for g, c in enumerate(owners)]]
    for _, s in lanes[0]:
        s.connect_local([x for _, x in lanes[0]])
    PeerShardedSearch.sync_keys_local([x for _, x in lanes[0]])
    for c in owners:
        c.sync()
    for l in range(1, n_lanes):
        cs = [liorf_b200.Context() for _ in range(world)]
        for g, c in enumerate(cs):
            c.scBorrowDatabase(owners[g])
        L = [(c, PeerShardedSearch(c, g, world, bounds, Q, torch, k_total_max=0)) for g, c in enumerate(cs)]
        for _, s in L:
            s.connect_local([x for _, x in L])
        lanes.append(L)
    # unsharded reference with the same number of lanes
    ref = liorf_b200.Context(); load(ref, 0, K)
    ref_lanes = [(ref, PeerShardedSearch(ref, 0, 1, [0, K], Q, torch))]
    for l in range(1, n_lanes):
        c = liorf_b200.Context(); c.scBorrowDatabase(ref)
        ref_lanes.append((c, PeerShardedSearch(c, 0, 1, [0, K], Q, torch)))
    for _, s in ref_lanes:
        s.connect_local([s])

    n_src = min(K, 2000)
    src_rows = (np.arange(n_src, dtype=np.int64) * K) // n_src          # loop sources spread over all rows (and so over all ranks)
    sample = np.concatenate([synth.sc_descriptors(1, first=int(i)) for i in src_rows])
    qd, src, shift = synth.sc_queries(sample, Q)
    src = np.where(src >= 0, src_rows[np.maximum(src, 0)], -1)
    d_q = torch.from_numpy(qd).to(dev)
    torch.cuda.synchronize()

    r_out = [t.cpu().numpy() for t in ref_lanes[0][1].query(d_q)]
    ref.sync()
    r_out = [t.cpu().numpy() for t in ref_lanes[0][1].query(d_q)]
    ref.sync()
    print("unsharded: planted found %d of %d" % (int(((r_out[0] == src) & (src >= 0)).sum()), int((src >= 0).sum())))

    # 1 + 2: complete step-wise batches (twice: the second has every buffer sized), per-step times of rank 0
    names = {1: "stage 1 of the slice (keys, images, GEMM, top3, select, re-rank + push C)", 2: "collect + stage 2 of the owned pairs + push D", 4: "decide"}
    for L in lanes:
        for rep in range(2):
            t = {}
            for step in (1, 2, 4):
                ev = []
                for g, (c, s) in enumerate(L):
                    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                    with torch.cuda.stream(s.stream):
                        e0.record()
                    res = s.query(d_q, phases=step)
                    with torch.cuda.stream(s.stream):
                        e1.record()
                    c.sync()                            # one rank at a time
                    ev.append((e0, e1))
                t[step] = [a.elapsed_time(b) for a, b in ev]
        for g, (c, s) in enumerate(L):
            out = [x.cpu().numpy() for x in s._out[Q]]
            same = all(np.array_equal(np.nan_to_num(a, nan=-7.0), np.nan_to_num(b, nan=-7.0)) for a, b in zip(out, r_out))
            if not same:
                print("MISMATCH vs unsharded on rank", g); sys.exit(1)
    print("all %d ranks x %d lanes bit-equal to the unsharded search" % (world, n_lanes))
    for step in (1, 2, 4):
        print("%-80s rank 0 %.3f ms   mean %.3f   max %.3f" % (names[step], t[step][0], float(np.mean(t[step])), float(np.max(t[step]))))
    print("sum over steps: rank 0 %.3f ms for %d queries, %d keys, %d ranks (ideal = unsharded / %d)" % (sum(t[k][0] for k in t), Q, K, world, world))

    if os.environ.get("LIORF_PROF_MODE") == "ncu":       # launch list of ONE rank's batch: run under `ncu --metrics gpu__time_duration.sum`; the last batch is the one to read
        c, s = lanes[0][0]
        c.lib.liorf_sc_shard_debug_nowait(c.h, 1)
        for _ in range(3):
            s.query(d_q, phases=1); s.query(d_q, phases=2); s.query(d_q, phases=4)
            c.sync()
        print("ncu mode: 3 batches of rank 0 issued (plain launches)")
        return
    # 3: throughput, rank 0 alone vs unsharded, same lanes in flight
    def throughput(Ls, tag):
        streams = [s.stream for _, s in Ls]
        for _ in range(4):
            for _, s in Ls:
                s.query(d_q)
        for c, _ in Ls:
            c.sync()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(streams[0]):
            e0.record()
        for st in streams[1:]:
            st.wait_event(e0)
        for _ in range(reps):
            for _, s in Ls:
                s.query(d_q)
        for st in streams[1:]:
            streams[0].wait_stream(st)
        with torch.cuda.stream(streams[0]):
            e1.record()
        for c, _ in Ls:
            c.sync()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (reps * len(Ls))
        print("%-34s %.4f ms per batch of %d queries (%d lanes in flight) = %.1f M queries/s" % (tag, ms, Q, len(Ls), Q / ms / 1e3))
        return ms
    t1 = throughput(ref_lanes, "unsharded (1 GPU):")
    r0 = [L[0] for L in lanes]
    for c, s in r0:
        c.lib.liorf_sc_shard_debug_nowait(c.h, 1)
    tg = throughput(r0, "rank 0 of %d alone (no waits):" % world)
    print("predicted speed-up of %d GPUs over one: %.2fx = %.2f of linear (peer waits not included)" % (world, t1 / tg, t1 / tg / world))
    one = [r0[0]]
    t1_1 = throughput([ref_lanes[0]], "unsharded, 1 lane:")
    tg_1 = throughput(one, "rank 0 of %d alone, 1 lane:" % world)
    print("one batch in flight: %.2fx = %.2f of linear" % (t1_1 / tg_1, t1_1 / tg_1 / world))
    for L in lanes[1:]:
        for c, _ in L:
            c.close()
    for c, _ in ref_lanes[1:]:
        c.close()
    for c in owners:
        c.close()
    ref.close()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Per-kernel view of ONE batch of the peer-window sharded ScanContext search, with the ranks emulated as contexts on one GPU
(stepwise enqueue).  Run under `ncu --metrics gpu__time_duration.sum --csv` for the launch list, or plainly for event times.
usage: python tools/profile_sc_shard.py [K] [Q] [world]"""
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
    world = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    kloc = K // world
    ctxs = [liorf_b200.Context() for _ in range(world)]
    for g, c in enumerate(ctxs):
        c.reserve(1024, 1024, 0, kloc)
        for s in range(0, kloc, 10000):
            c.scAddDescriptors(synth.sc_descriptors(min(10000, kloc - s), first=g * kloc + s))
    S = [PeerShardedSearch(c, g, world, [r * kloc for r in range(world + 1)], Q, torch) for g, c in enumerate(ctxs)]
    for s in S:
        s.connect_local(S)
    n_src = min(K, 2000)
    src_rows = (np.arange(n_src, dtype=np.int64) * K) // n_src          # loop sources spread over all rows (and so over all ranks)
    sample = np.concatenate([synth.sc_descriptors(1, first=int(i)) for i in src_rows])
    qd, src, shift = synth.sc_queries(sample, Q)
    src = np.where(src >= 0, src_rows[np.maximum(src, 0)], -1)
    dq = []
    for s in S:
        with torch.cuda.stream(s.stream):
            dq.append(torch.from_numpy(qd).to(s.dev))
    names = {16: "keys of the own query slice + push K", 1: "gather keys + images + GEMM + top3 + push T", 2: "threshold + select + re-rank + push C", 4: "merge + distance + push D", 8: "decide"}
    for rep in range(3):
        t = {}
        for step in (16, 1, 2, 4, 8):
            ev = []
            for g, s in enumerate(S):
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(s.stream):
                    e0.record()
                res = s.query(dq[g], phases=step)
                with torch.cuda.stream(s.stream):
                    e1.record()
                ctxs[g].sync()                          # one rank at a time: its kernels have the GPU to themselves, as on its own device
                ev.append((e0, e1))
            t[step] = [a.elapsed_time(b) for a, b in ev]
        if rep == 2:
            for step in (16, 1, 2, 4, 8):
                print("%-46s rank 0 %.3f ms   mean %.3f   max %.3f" % (names[step], t[step][0], float(np.mean(t[step])), float(np.max(t[step]))))
            print("sum over steps: rank 0 %.3f ms, sum of the per-step maxima %.3f ms, for %d queries, %d keys per rank, %d ranks" %
                  (sum(t[k][0] for k in t), sum(max(t[k]) for k in t), Q, kloc, world))
    loop = res[0].cpu().numpy()
    print("planted found:", int(((loop == src) & (src >= 0)).sum()), "of", int((src >= 0).sum()))
    for c in ctxs:
        c.close()


if __name__ == "__main__":
    main()

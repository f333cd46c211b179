#!/usr/bin/env python
"""Times the ScanContext ring-key search (config 5 shape) on one GPU: CUDA-core brute force vs the tcgen05 filter.
usage: python tools/profile_sc.py [K] [Q] [reps]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import liorf_b200
    from tools import synth
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    Q = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    ctx = liorf_b200.Context()
    ctx.reserve(1024, 1024, 0, K)
    for s in range(0, K, 10000):
        ctx.scAddDescriptors(synth.sc_descriptors(min(10000, K - s), first=s))
    sample = synth.sc_descriptors(min(K, 2000), first=0)
    qd, src, shift = synth.sc_queries(sample, Q)
    dev = torch.device("cuda:0")
    st = torch.cuda.ExternalStream(ctx.stream(), device=dev)
    with torch.cuda.stream(st):
        d_q = torch.from_numpy(qd).to(dev)
        keys = torch.empty((Q, 20), dtype=torch.float32, device=dev); sk = torch.empty((Q, 60), dtype=torch.float64, device=dev); cn = torch.empty_like(sk)
        ctx.lib.liorf_sc_prepare_queries_dev(ctx.h, C.c_void_p(d_q.data_ptr()), Q, C.c_void_p(keys.data_ptr()), None, None)
        d = torch.empty((Q, 3), dtype=torch.float32, device=dev); i = torch.empty((Q, 3), dtype=torch.int32, device=dev)
        pd = torch.empty((Q, 3), dtype=torch.float64, device=dev); ps = torch.empty((Q, 3), dtype=torch.int32, device=dev)
    out = {}
    for mode, name in ((1, "brute"), (2, "tensor")):
        ctx.scSetSearchPath(mode)
        for _ in range(3):
            ctx.lib.liorf_sc_knn_batch_dev(ctx.h, C.c_void_p(keys.data_ptr()), Q, 0, C.c_void_p(d.data_ptr()), C.c_void_p(i.data_ptr()))
        ctx.sync()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            e0.record()
        for _ in range(reps):
            ctx.lib.liorf_sc_knn_batch_dev(ctx.h, C.c_void_p(keys.data_ptr()), Q, 0, C.c_void_p(d.data_ptr()), C.c_void_p(i.data_ptr()))
        with torch.cuda.stream(st):
            e1.record()
        ctx.sync()
        out[name] = (e0.elapsed_time(e1) / reps, d.cpu().numpy().copy(), i.cpu().numpy().copy())
        print("%s: %.3f ms per batch of %d queries over %d keys (%.2f M queries/s)" % (name, out[name][0], Q, K, Q / out[name][0] / 1e3))
    print("tensor stats:", ctx.scTensorStats(), "per query: %.1f" % (ctx.scTensorStats()["candidates"] / Q))
    print("identical results:", np.array_equal(out["brute"][2], out["tensor"][2]) and np.array_equal(out["brute"][1].view(np.uint32), out["tensor"][1].view(np.uint32)))
    # stage 2 alone
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    for rep in range(2):
        with torch.cuda.stream(st):
            e0.record()
        for _ in range(reps):
            ctx.lib.liorf_sc_distance_batch_dev(ctx.h, C.c_void_p(d_q.data_ptr()), C.c_void_p(i.data_ptr()), Q, 0, C.c_void_p(pd.data_ptr()), C.c_void_p(ps.data_ptr()))
        with torch.cuda.stream(st):
            e1.record()
        ctx.sync()
    print("stage 2 (distanceBtnScanContext, %d pairs): %.3f ms" % (3 * Q, e0.elapsed_time(e1) / reps))
    with torch.cuda.stream(st):
        e0.record()
    for _ in range(reps):
        ctx.lib.liorf_sc_prepare_queries_dev(ctx.h, C.c_void_p(d_q.data_ptr()), Q, C.c_void_p(keys.data_ptr()), None, None)
    with torch.cuda.stream(st):
        e1.record()
    ctx.sync()
    print("prepare queries (keys from descriptors): %.3f ms" % (e0.elapsed_time(e1) / reps))
    ctx.close()


if __name__ == "__main__":
    main()

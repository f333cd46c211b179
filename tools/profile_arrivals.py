"""%globaltimer timeline of the persistent solver: worker arrival spread, reducer latency, flag propagation.
usage: python tools/profile_arrivals.py [preroll] [single]   (single: config-1 style 30 forced iterations on the last frame)"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def report(ctx, iters, W):
    buf = (C.c_ulonglong * (64 * 160))()
    ctx.lib.liorf_debug_s2m_arrivals(ctx.h, 1, buf)
    g = np.array(list(buf), np.int64).reshape(64, 160)
    prev_pub = None
    for k in range(min(iters, 64)):
        arr = g[k, :W]; ready, pub, seen = g[k, 156], g[k, 157], g[k, 158]
        base = prev_pub if prev_pub is not None else arr.min()
        a = np.sort(arr - base) / 1e3
        print(f"   iter {k:2d}: arrivals after previous publish: first {a[0]:6.1f}  p50 {a[len(a) // 2]:6.1f}  p90 {a[int(len(a) * 0.9)]:6.1f}  last {a[-1]:6.1f} us;"
              f"  last arrival -> sums ready {(ready - arr.max()) / 1e3:5.1f};  solve+publish {(pub - ready) / 1e3:5.1f};  publish -> worker 0 sees flag {(seen - pub) / 1e3:5.1f};"
              f"  iteration {((pub - base) / 1e3):6.1f} us")
        prev_pub = pub


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    single = len(sys.argv) > 2
    seq = bench.Sequence(P + 8, 0)
    for i in range(P + 8):
        seq.frame(i)
    pipe = bench.GpuPipeline(seq, 0)
    pipe.stage(range(P + 8))
    for i in range(P):
        pipe.step(i, "dev")
    ctx = pipe.ctx
    ctx.lib.liorf_debug_s2m_arrivals(ctx.h, 1, None)
    W = 131
    for i in range(P, P + 3):
        pipe.step(i, "dev")
        c = ctx.lastCounts()
        print(f"frame {i}: n_ds={c['n_ds']} iters={c['iters']}")
        report(ctx, c["iters"], W)
    if single:
        import torch
        raw, (t0, it, rot, ptr) = seq.frame(P + 2)
        xyz = np.stack([raw["x"], raw["y"], raw["z"], raw["i"]], 1).astype(np.float32)
        ids = ctx.extractNearby(t0, 2.0)
        guess = (pipe.prev + np.array([np.deg2rad(0.5), np.deg2rad(0.3), np.deg2rad(1.5), 0.35, 0.1, 0.02], np.float32)).astype(np.float32)
        ctx.setCurrentScan(xyz)
        ctx.extractSurroundingKeyFrames(ids, want_count=False)
        ctx.downsampleCurrentScan(want_output=False)
        ctx.scan2MapOptimizationAsync(guess, 30, True)
        ctx.getPose()
        c = ctx.lastCounts()
        print(f"single frame: n_ds={c['n_ds']} m_ds={c['m_ds']} iters={c['iters']}")
        report(ctx, 12, W)
    ctx.close()


if __name__ == "__main__":
    main()

"""Phase clocks of the persistent solver on frames of the sequence workload (small scans: ~4.3k queries, ~3 iterations).
usage: python tools/profile_seq_solver.py [preroll]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    seq = bench.Sequence(P + 12, 0)
    for i in range(P + 12):
        seq.frame(i)
    pipe = bench.GpuPipeline(seq, 0)
    pipe.stage(range(P + 12))
    for i in range(P):
        pipe.step(i, "dev")
    ctx = pipe.ctx
    ctx.lib.liorf_debug_s2m_clocks(ctx.h, 1, None)
    names = ["loop(knn+fit)", "block reduce", "arrive", "wait for flag", "read result"]
    for i in range(P, P + 6):
        ctx.enableTiming(True)
        pipe.step(i, "dev")
        tm = ctx.getTiming()
        buf = (C.c_longlong * 512)()
        ctx.lib.liorf_debug_s2m_clocks(ctx.h, 1, buf)
        full = np.array(list(buf), np.int64).reshape(64, 8)
        c = ctx.lastCounts()
        it = c["iters"]
        d = full[:it, :6]
        ph = np.diff(d, axis=1)
        print(f"frame {i}: n_ds={c['n_ds']} iters={it} solver section {tm['scan2map'][0] / max(tm['scan2map'][1], 1) * 1e3:.1f} us; kernel span (first stamp to last) {(d[-1, 5] - d[0, 0]) / 1.965e3:.1f} us")
        for k in range(it):
            print("   iter %d: " % k + "  ".join(f"{n} {ph[k, j] / 1.965e3:6.1f}us" for j, n in enumerate(names)) + f"   last-CTA sum {full[k, 6] / 1.965e3:.1f}us solve {full[k, 7] / 1.965e3:.1f}us")
    ctx.close()


if __name__ == "__main__":
    main()

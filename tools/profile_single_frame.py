"""Config 1 (kitti64_single) driver for timing experiments and ncu captures:
50 keyframes at 1 m spacing (each = HDL-64 scan voxelised at 0.4 m), local map at 0.5 m, one ~119k-point query scan,
downsample + grid build + 30 forced LM iterations.   python tools/profile_single_frame.py [reps] [n_keyframes]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import liorf_b200  # noqa: E402
from tools import synth  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    nkf = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    ctx = liorf_b200.Context(downsampleRate=1, point_filter_num=1)
    t = time.time()
    for k in range(nkf):
        pose = np.array([0, 0, 0, 1.0 * k, 0, 0], np.float64)
        xyz = synth.raw_to_xyzi(synth.scan(synth.HDL64, pose, seed=synth.SEED0 + k))
        ctx.setCurrentScan(xyz)
        ctx.downsampleCurrentScan(want_output=False)
        ctx.addKeyframe(pose.astype(np.float32), 0.1 * k)
    qpose = np.array([0, 0, 0, 1.0 * (nkf - 1), 0, 0], np.float64)
    scan = synth.raw_to_xyzi(synth.scan(synth.HDL64, qpose, seed=synth.SEED0 + 500))
    init = (qpose + np.array([np.deg2rad(0.5), np.deg2rad(0.3), np.deg2rad(1.5), 0.35, 0.1, 0.02])).astype(np.float32)
    print(f"setup {time.time() - t:.1f}s; scan {len(scan)} pts", flush=True)
    ids = list(range(nkf))
    kf0 = ctx.getKeyframe(0)[1]
    walls = []
    for r in range(reps + 3):
        if r == 3:
            ctx.enableTiming(True)                     # after the warm-up repetitions (first calls allocate)
            walls = []
        ctx.updateKeyframePose(0, kf0)
        ctx.setCurrentScan(scan)
        ctx.sync()
        t0 = time.perf_counter()
        ctx.extractSurroundingKeyFrames(ids, want_count=False)
        ctx.downsampleCurrentScan(want_output=False)
        ctx.scan2MapOptimizationAsync(init, 30, True)
        pose, tr = ctx.getPose(want_trace=True)
        walls.append(time.perf_counter() - t0)
    c = ctx.lastCounts()
    tm = ctx.getTiming()
    print("counts", c, "final pose", pose, "iters", tr.iters, "nsel", tr.nsels()[-1])
    for k, (ms, calls) in tm.items():
        if calls:
            print(f"  {k:12s} {ms / calls * 1e3:9.1f} us/call  ({calls} calls)")
    print(f"wall (map build + downsample + solve) median {np.median(walls) * 1e3:.3f} ms")
    print(f"solver: {tm['scan2map'][0] / tm['scan2map'][1] / 30 * 1e3:.2f} us/iteration")
    import ctypes as C
    ctx.lib.liorf_debug_s2m_clocks(ctx.h, 1, None)
    ctx.scan2MapOptimizationAsync(init, 30, True); ctx.getPose()
    buf = (C.c_longlong * 512)()
    ctx.lib.liorf_debug_s2m_clocks(ctx.h, 1, buf)
    full = np.array(list(buf), np.int64).reshape(64, 8)[:30]
    d = full[:, :6]
    ph = np.diff(d, axis=1)
    print(f"last CTA: sum partials {np.median(full[1:, 6]):.0f} cyc, solve+publish {np.median(full[1:, 7]):.0f} cyc (iter0: {full[0, 6]}, {full[0, 7]})")
    names = ["loop(knn+fit)", "block reduce", "arrive", "wait for flag", "read result"]
    print("phase cycles (median over iterations 1..29), CTA 0:")
    for k, nme in enumerate(names):
        print(f"   {nme:14s} {np.median(ph[1:, k]):9.0f} cyc   iter0 {ph[0, k]:9.0f}")
    print(f"   per-iteration total {np.median(d[2:, 0] - d[1:-1, 0]):9.0f} cyc")
    ctx.close()


if __name__ == "__main__":
    main()

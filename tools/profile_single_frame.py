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
    d = full[:, [0, 3, 1, 2, 4, 5]]              # start, lists rebuilt (phases 1 + 2), rows done (phase 3), partial published, answer seen, broadcast
    ph = np.diff(d, axis=1)
    print(f"last CTA: sum partials {np.median(full[1:, 6]):.0f} cyc, solve+publish {np.median(full[1:, 7]):.0f} cyc (iter0: {full[0, 6]}, {full[0, 7]})")
    names = ["check + rebuild", "5-NN + rows", "products", "wait for answer", "broadcast"]
    print("phase cycles (median over iterations 1..29), CTA 0:")
    for k, nme in enumerate(names):
        print(f"   {nme:14s} {np.median(ph[1:, k]):9.0f} cyc   iter0 {ph[0, k]:9.0f}")
    print(f"   per-iteration total {np.median(d[2:, 0] - d[1:-1, 0]):9.0f} cyc")
    for k in range(6):
        print(f"   iteration {k}: " + "  ".join(f"{nme} {ph[k, j]}" for j, nme in enumerate(names)))
    # %globaltimer timeline: arrival spread of the workers, reducer latency, flag propagation
    ctx.lib.liorf_debug_s2m_clocks(ctx.h, 0, None)
    ctx.lib.liorf_debug_s2m_arrivals(ctx.h, 1, None)
    ctx.scan2MapOptimizationAsync(init, 30, True); ctx.getPose()
    gb = (C.c_ulonglong * (64 * 160))()
    ctx.lib.liorf_debug_s2m_arrivals(ctx.h, 1, gb)
    g = np.array(list(gb), np.int64).reshape(64, 160)
    W = int(np.count_nonzero(g[1, :156]))
    ev = g[40:44, :W]
    print("   events over iterations 10..29 (all workers): list rebuilds %d, ordered insert passes %d, plane refits %d, overflowed lists %d" % tuple(ev.sum(1)))
    for nm, row in zip(["rebuilds", "insert passes", "refits", "overflowed lists"], ev):
        top = np.argsort(row)[-3:][::-1]
        print(f"      most {nm}: " + ", ".join(f"worker {int(w)}: {int(row[w])}" for w in top))
    prev_pub = None
    for k in range(30):
        arr = g[k, :W]; ready, pub, seen = g[k, 156], g[k, 157], g[k, 158]
        base = prev_pub if prev_pub is not None else arr.min()
        a = np.sort(arr - base) / 1e3
        if k < 8 or k % 5 == 0:
            print(f"   iter {k:2d}: arrivals after previous publish: first {a[0]:6.1f}  p50 {a[len(a) // 2]:6.1f}  p90 {a[int(len(a) * 0.9)]:6.1f}  last {a[-1]:6.1f} us;"
                  f"  last arrival -> sums ready {(ready - arr.max()) / 1e3:5.1f};  solve+publish {(pub - ready) / 1e3:5.1f};  publish -> worker 0 sees it {(seen - pub) / 1e3:5.1f};"
                  f"  iteration {((pub - base) / 1e3):6.1f} us")
        if k == 10:
            late = np.argsort(arr)[-5:]
            print("      latest workers of iteration 10:", [(int(w), round(float(arr[w] - base) / 1e3, 1)) for w in late])
        prev_pub = pub
    # phase stamps of the one-kernel VoxelGrid (scan side)
    ctx.lib.liorf_debug_voxelgrid_stamps(ctx.h, 1, None)
    ctx.setCurrentScan(scan); ctx.downsampleCurrentScan(want_output=False); ctx.sync()
    ctx.setCurrentScan(scan); ctx.downsampleCurrentScan(want_output=False); ctx.sync()
    vb = (C.c_ulonglong * 32)()
    ctx.lib.liorf_debug_voxelgrid_stamps(ctx.h, 1, vb)
    v = np.array(list(vb), np.int64); v = v[v > 0]
    print("VoxelGrid (one kernel) phase stamps, us since kernel start:", [round(float(x - v[0]) / 1e3, 1) for x in v])
    ctx.close()


if __name__ == "__main__":
    main()

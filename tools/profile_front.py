#!/usr/bin/env python
"""Streaming front end on the KITTI-shaped scan, for ncu captures of the bandwidth-side kernels (north star: "achieved HBM/L2 GB/s
against ~8 TB/s for the streaming and gather kernels"): projectPointCloud (deskew, yaml filters off so that all ~119k points are kept),
downsampleCurrentScan, extractSurroundingKeyFrames (transform + concat + VoxelGrid + grid build over ~670k raw map points).
usage: python tools/profile_front.py [reps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import liorf_b200
    from tools import synth
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    ctx = liorf_b200.Context(downsampleRate=1, point_filter_num=1)
    ctx.reserve(1 << 18, 4 << 20, 4 << 20, 0)
    nkf = 50
    for k in range(nkf):
        pose = np.array([0, 0, 0, 1.0 * k, 0, 0], np.float64)
        ctx.setCurrentScan(synth.raw_to_xyzi(synth.scan(synth.HDL64, pose, seed=synth.SEED0 + k)))
        ctx.downsampleCurrentScan(want_output=False)
        ctx.addKeyframe(pose.astype(np.float32), 0.1 * k)
    omega = (0.01, -0.02, 0.3)
    raw = synth.scan(synth.HDL64, (0, 0, 0, 49.0, 0, 0), omega=omega, seed=synth.SEED0 + 500)
    it, rot, ptr = synth.imu_table(100.0, 100.0 + float(raw["time"][-1]), omega, rate_hz=100.0)
    kf0 = ctx.getKeyframe(0)[1]
    ctx.enableTiming(True)
    for r in range(reps):
        ctx.projectPointCloud(raw, 100.0, it, rot, ptr, True, want_output=False)
        ctx.downsampleCurrentScan(want_output=False)
        ctx.updateKeyframePose(0, kf0)
        ctx.extractSurroundingKeyFrames(list(range(nkf)), want_count=False)
        ctx.sync()
    tm = ctx.getTiming()
    c = ctx.lastCounts()
    print("raw %d points; sections (us/call):" % len(raw), {k: round(v[0] / max(v[1], 1) * 1e3, 1) for k, v in tm.items() if v[1]})
    ctx.close()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Small-shape pass over every kernel family of the library, meant to run under compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
Shapes are tiny (the tools slow kernels down 10-100x); results are still checked against the oracle where that is cheap."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)


def main():
    import torch
    import liorf_b200
    import pyoracle as orc
    from tools import synth
    from liorf_b200.sc_sharded import PeerShardedSearch
    rng = np.random.default_rng(1)
    # deskew + VoxelGrid (single-CTA and multi-kernel paths) + local map + grid + persistent solver
    kfs = []
    for k in range(3):
        pose = np.array([0, 0, 0, 1.0 * k, 0, 0], np.float32)
        ds, _, _ = orc.voxel_grid(synth.raw_to_xyzi(synth.scan(synth.HDL64, pose.astype(np.float64), seed=synth.SEED0 + k))[::4], 0.4)
        kfs.append((ds, pose))
    c = liorf_b200.Context()
    for cl, p in kfs:
        c.addKeyframeCloud(cl, p)
    m = c.extractSurroundingKeyFrames([0, 1, 2])
    raw = synth.scan(synth.HDL64, (0, 0, 0, 2.0, 0, 0), omega=(0.01, 0.0, 0.3), seed=synth.SEED0 + 9)
    it, rot, ptr = synth.imu_table(100.0, 100.0 + float(raw["time"][-1]), (0.01, 0.0, 0.3), rate_hz=100.0)
    out, n_kept = c.projectPointCloud(raw, 100.0, it, rot, ptr, True)
    o_out, _ = orc.project_point_cloud(raw, dict(lidarMinRange=1.0, lidarMaxRange=1000.0, N_SCAN=64, downsampleRate=2, point_filter_num=5), 100.0, it, rot, ptr, True)
    assert n_kept == len(o_out)
    ds, n = c.downsampleCurrentScan(len(raw))
    c.forceLargeVoxelGrid(True); ds2, n2 = c.downsampleCurrentScan(len(raw)); c.forceLargeVoxelGrid(False)
    assert n == n2 and np.array_equal(ds, ds2)
    pose, tr = c.scan2MapOptimization(np.array([0.002, -0.001, 0.01, 2.1, 0.05, 0.0], np.float32), 6, force_all_iters=True)
    assert tr.iters == 6
    g = c.surfOptimization(pose, n); c.combineOptimizationCoeffs(n); c.LMOptimization(0, pose)
    # ScanContext: make + live detect, batched search on both ring-key paths, two ranks through peer windows on one device
    c.makeAndSaveScancontextAndKeys(out)
    db = synth.sc_descriptors(700, seed=3)
    qd, src, _ = synth.sc_queries(db, 96, seed=4)
    s = liorf_b200.Context(); s.scAddDescriptors(db)
    s.scSetSearchPath(1); a = s.scQueryBatch(qd)
    s.scSetSearchPath(2); b = s.scQueryBatch(qd)
    assert all(np.array_equal(np.nan_to_num(x, nan=-7.0), np.nan_to_num(y, nan=-7.0)) for x, y in zip(a, b))
    for _ in range(3):
        s.detectLoopClosureID()
    ctxs = [liorf_b200.Context(), liorf_b200.Context()]
    ctxs[0].scAddDescriptors(db[:300]); ctxs[1].scAddDescriptors(db[300:])
    for cc in ctxs:
        cc.scSetSearchPath(2)
    S = [PeerShardedSearch(cc, g_, 2, [0, 300, 700], 96, torch) for g_, cc in enumerate(ctxs)]
    for x in S:
        x.connect_local(S)
    PeerShardedSearch.sync_keys_local(S)
    dq = torch.from_numpy(qd).cuda()
    torch.cuda.synchronize()
    for rep in range(2):
        for step in (1, 2, 4):
            res = [x.query(dq, phases=step) for x in S]
        for cc in ctxs:
            cc.sync()
    for r in res:
        assert np.array_equal(r[0].cpu().numpy(), a[0]) and np.array_equal(r[3].cpu().numpy(), a[3])
    # loop-closure ICP + global map on the three keyframes
    r = c.loopClosureICP(2, 0, history_search_num=1, loop_index=-1, icp_leaf=0.5, max_corr_dist=5.0, max_iters=5)
    c.buildGlobalMap(1000.0, 1.0, 1.0)
    for cc in ctxs + [s, c]:
        cc.close()
    print("sanitize_smoke: ok (N_ds=%d, M=%d, icp ran=%d)" % (n, m, r.ran))


if __name__ == "__main__":
    main()

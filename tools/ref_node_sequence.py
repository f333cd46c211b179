"""config 2 on the CPU, twice: the reference's OWN mapOptimization node (src/mapOptmization.cpp compiled unchanged, oracle/_ref/libliorf_ref_mapopt.so),
one liorf::cloud_info per scan through laserCloudInfoHandler, against bench_common.CpuPipeline — the oracle pipeline tools/run_kitti05_full.py holds the
GPU's 300-frame prefix to (profiles/r02_kitti05_full.json: 0 decision mismatches).  Each side follows its own pose chain from frame 0 with the same
guess function (previous pose o true increment, perturbed).  TEST INFRASTRUCTURE; no GPU.

usage: python tools/ref_node_sequence.py [n_frames=300] > profiles/r02_reference_node_vs_cpu_pipeline.txt"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench_common as bc      # noqa: E402
import pyoracle as o           # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    seq = bc.Sequence(n + 1, 0, loop=True)                          # the drive of tools/run_kitti05_full.py
    cpu = bc.CpuPipeline(seq)
    ref = o.RefMapOpt(useImuHeadingInitialization=1)
    prev = None
    mism = equal = 0
    dt = dr = 0.0
    t0w = time.time()
    for i in range(n):
        raw, (t0, it, rot, ptr) = seq.frame(i)
        cloud, _ = o.project_point_cloud(raw, seq.filters, t0, it, rot, ptr, True)
        g = np.asarray(seq.initial_guess(i, prev), np.float32)
        ref.set_transform(g)                                        # transformTobeMapped = this frame's guess; no odometry / IMU increment on top (:922, :946)
        ref.cloud_info(t0, cloud, 0, 0, g[:3], None)
        rs = ref.state(); prev = rs["tf"].copy()
        p = cpu.step(i)
        equal += int(np.array_equal(p.view(np.uint32), rs["tf"].view(np.uint32)))
        mism += int(len(cpu.kf_clouds) != rs["keyframes"])
        dt = max(dt, float(np.abs(p[3:] - rs["tf"][3:]).max())); dr = max(dr, float(np.abs(p[:3] - rs["tf"][:3]).max()))
        if i % 50 == 0:
            print("frame %4d  pose %s  keyframes %d / %d  n_ds %d  m_ds %d" % (i, "bit-equal" if np.array_equal(p.view(np.uint32), rs["tf"].view(np.uint32)) else "DIFFERENT",
                                                                           len(cpu.kf_clouds), rs["keyframes"], rs["n_ds"], rs["m_ds"]))
    print("frames %d: pose bit-equal after %d of them, keyframe-decision mismatches %d, max |dt| %.3g m, max |dr| %.3g rad, keyframes %d, ScanContext entries %d (%.0f s)"
          % (n, equal, mism, dt, dr, rs["keyframes"], rs["sc_entries"], time.time() - t0w))
    return 0 if equal == n and mism == 0 else 1


if __name__ == "__main__":
    sys.exit(main())

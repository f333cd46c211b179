#!/usr/bin/env python
"""BASELINE config 2 AT LENGTH: a synthetic KITTI-05-length drive (2 761 scans at 10 Hz, 64 beams, closed rectangle that revisits its first
leg → real loop closures) end to end through liorf_process_frame on one GPU — deskew, registration, keyframe gate, keyframe store,
ScanContext make, detectLoopClosureID every 10th scan — plus the oracle's CPU pipeline on a prefix, compared frame by frame.
usage: python tools/run_kitti05_full.py [n_scans=2761] [oracle_prefix=300] [out.json]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2761
    n_orc = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    out_path = sys.argv[3] if len(sys.argv) > 3 else None
    seq = bench.Sequence(n + 1, 0, loop=True)
    gpu = bench.GpuPipeline(seq, 0)
    CH = 128                                                       # scans are synthesised and staged in chunks (pinned + device copies of a chunk: ~0.7 GB)
    cpu = bench.CpuPipeline(seq) if n_orc > 0 else None
    drift_t = drift_r = 0.0
    mism = dict(n_kept=0, n_ds=0, keyframe=0, iters=0)
    loops, kf_at = [], []
    gpu_ms = 0.0; wall = 0.0; t_synth = 0.0
    ext = torch.cuda.ExternalStream(gpu.ctx.stream(), device=torch.device("cuda:0"))
    for c0 in range(0, n, CH):
        c1 = min(n, c0 + CH)
        a = time.perf_counter()
        for i in range(c0, min(n + 1, c1 + 1)):
            seq.frame(i)
        gpu.stage(range(c0, min(n + 1, c1 + 1)))
        t_synth += time.perf_counter() - a
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        with torch.cuda.stream(ext):
            e0.record()
        outs = []
        for i in range(c0, c1):
            guess = seq.initial_guess(i, gpu.prev)
            nxt = gpu._frame_in(i + 1, "e2e") if (i + 1) in gpu.pin_raw else None
            fo = gpu.ctx.processFrameIn(gpu._frame_in(i, "e2e"), guess, nxt)       # HOST (pinned) scans: H2D inside the timed region
            gpu.prev = np.array(fo.pose[:], np.float32)
            outs.append((i, gpu.prev.copy(), fo.n_kept, fo.n_ds, fo.m_ds, fo.iters, fo.is_keyframe, fo.loop_checked, fo.loop_id, fo.loop_yaw))
        with torch.cuda.stream(ext):
            e1.record()
        torch.cuda.synchronize()
        wall += time.perf_counter() - w0
        gpu_ms += e0.elapsed_time(e1)
        for (i, pose, nk, nds, mds, its, kf, lc, lid, lyaw) in outs:
            if kf:
                kf_at.append(i)
            if lc and lid >= 0:
                loops.append((i, int(lid), float(lyaw)))
            if cpu is not None and i < n_orc:                      # the CPU pipeline on the same frames, from ITS OWN pose chain (no re-synchronisation)
                nk0 = len(cpu.kf_clouds)
                o_pose = cpu.step(i)
                drift_t = max(drift_t, float(np.max(np.abs(pose[3:] - o_pose[3:])))); drift_r = max(drift_r, float(np.max(np.abs(pose[:3] - o_pose[:3]))))
                mism["keyframe"] += int(kf != (len(cpu.kf_clouds) - nk0))
                if "ds" in getattr(cpu, "last", {}):
                    mism["n_ds"] += int(nds != len(cpu.last["ds"])); mism["iters"] += int(its != cpu.last.get("iters", its))
        for i in range(c0, c1):                                    # free the chunk (the look-ahead frame c1 stays)
            for d in (gpu.dev_raw, gpu.pin_raw, seq.raw, seq.imu):
                d.pop(i, None)
            gpu.fin.pop((i, "dev"), None); gpu.fin.pop((i, "e2e"), None)
        print(f"frames {c0}..{c1 - 1}: keyframes {len(kf_at)}, loops {len(loops)}, {gpu_ms / c1:.4f} ms/frame so far", file=sys.stderr, flush=True)
    nkf = gpu.ctx.numKeyframes()
    arena_pts = sum(gpu.ctx.getKeyframe(k)[0].shape[0] for k in range(0, nkf, max(1, nkf // 50))) / max(1, len(range(0, nkf, max(1, nkf // 50)))) * nkf
    truth = seq.poses[n - 1]
    res = dict(workload="kitti05_seq at length: %d scans (10 Hz, 64 beams, ~119k returns, closed rectangle with revisits), yaml filters, leaf 0.4 / 0.5, early-exit LM, "
                        "liorf_process_frame with look-ahead from pinned HOST scans, detectLoopClosureID every 10th scan" % n,
               scans=n, total_s_device=gpu_ms / 1e3, ms_per_frame_e2e=gpu_ms / n, wall_ms_per_frame=wall * 1e3 / n, keyframes=nkf, loops_found=len(loops),
               first_loops=loops[:8], keyframe_arena_bytes_estimate=int(arena_pts * 16), scancontext_db_entries=nkf, scancontext_db_bytes=nkf * 10560,
               final_pose=[float(v) for v in gpu.prev], final_truth=[float(v) for v in truth],
               final_position_error_m=float(np.linalg.norm(gpu.prev[3:] - truth[3:])), synth_and_staging_s=t_synth,
               oracle_prefix=dict(frames=min(n_orc, n), max_abs_translation_diff_m=drift_t, max_abs_rotation_diff_rad=drift_r, decision_mismatches=mism,
                                  note="GPU and CPU pipelines each follow their own pose chain from frame 0 (no re-synchronisation): the difference is the accumulated drift between the two"))
    print(json.dumps(res))
    if out_path:
        json.dump(res, open(out_path, "w"), indent=1)
    gpu.ctx.close()


if __name__ == "__main__":
    main()

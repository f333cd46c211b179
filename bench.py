#!/usr/bin/env python
"""bench.py — headline benchmark of the liorf scan-to-map hot path on B200.

N = 1 (`python bench.py [--gpus 1]`) — headline = BASELINE.json configs[0] `kitti64_single`, the configuration the north star's
"< 1 ms" target is quoted on (SURVEY §8(d) cfg 1): a ~119k-return synthetic HDL-64 scan, filters off, registered against the local
map of 50 full-density keyframes at 1 m spacing (each the scan at that pose voxelised at 0.4 m; map leaf 0.5 m), 30 FORCED LM iterations.
  one STEP = downsampleCurrentScan (VoxelGrid of the 119k points) + kdtreeSurfFromMap->setInputCloud (here: the voxel-hash grid build
             over the resident map; the reference rebuilds its kd-tree inside every scan2MapOptimization call, src/mapOptmization.cpp:1302)
             + scan2MapOptimization (30 x {surfOptimization, combineOptimizationCoeffs, LMOptimization}) — the "scan-to-map solve".
  `value`  : ms/frame with the scan resident in HBM, CUDA events on the library's stream, L2 flushed before every timed step.
  `e2e`    : the same step through the C ABI from a HOST (pinned) scan: H2D of the 1.9 MB scan and D2H of the pose inside the region.
  extras   : `rows` (yaml-filter row, early-exit row, map build, configs 3 `os1_128_dense` and 4 `livox_deskew`), `sequence`
             (configs[1] `kitti05_seq`: a drive through liorf_process_frame), `sc` (config 5 on one GPU), `roofline`, `cpu_baseline`, `clocks`.
N > 1 (torchrun, one rank per GPU) — headline = BASELINE.json configs[4] `sc_100k`, the one workload of the path that shards:
  ScanContext loop-closure search over a 100 000-keyframe database sharded over the ranks, STRONG scaling; one STEP = one batch of
  `--sc-q` (32768) replicated queries answered by all ranks together (csrc/sc_shard.cuh: NVLink peer-window exchange from the kernels);
  `value` = queries/s; every rank checks its answers against the unsharded search of the same queries (rank 0 holds a full copy of
  the database for that) and the run fails if a single bit differs.  The same search on ONE GPU with the same number of batches in
  flight is printed at N = 1 as `sc` and at N > 1 as `sc.unsharded_same_run`.

`--impl reference` times the CPU path (oracle restatement + the reference's vendored nanoflann kd-trees from oracle/_ref) on the same
inputs: N = 1 the kitti64_single step, N > 1 the batched ScanContext search (rank 0 only), all host threads.
"""
import os

# every stream needs its own hardware channel: several ScanContext lanes per GPU wait on flags written by other GPUs (csrc/sc_shard.cuh);
# read by the CUDA driver when the context is created, so it is set before torch is imported
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import argparse  # noqa: E402
import hashlib  # noqa: E402
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from bench_common import (ORACLE_DIR, T0, DT, KITTI, Sequence, GpuPipeline, CpuPipeline, ClockSampler, load_peaks, dist_env,  # noqa: E402,F401
                          bench_batched, pose_to_T, T_to_pose)

PERTURB = np.array([np.deg2rad(0.5), np.deg2rad(0.3), np.deg2rad(1.5), 0.35, 0.1, 0.02])     # SURVEY §8(d) cfg 1 initial-pose error
SINGLE_CFGS = {
    # name: (synth sensor, N_SCAN, scan leaf, map leaf)
    "kitti64_single": ("HDL64", 64, 0.4, 0.5),
    "os1_128_dense": ("OS1_128", 128, 0.2, 0.5),
}


# ----------------------------------------------------------------------------------------------------------------
# config 1 / 3 inputs: seeded, identical bytes for the GPU arm and the CPU arm (sha256 printed by both)
# ----------------------------------------------------------------------------------------------------------------
def make_single_inputs(name, nkf=50):
    from tools import synth
    sensor = getattr(synth, SINGLE_CFGS[name][0])
    poses = [np.array([0, 0, 0, 1.0 * k, 0, 0], np.float64) for k in range(nkf)]
    raws = [synth.scan(sensor, p, seed=synth.SEED0 + k) for k, p in enumerate(poses)]
    qraw = synth.scan(sensor, poses[-1], seed=synth.SEED0 + 500)
    h = hashlib.sha256()
    for r in raws + [qraw]:
        h.update(r.tobytes())
    return dict(name=name, poses=poses, scans=[synth.raw_to_xyzi(r) for r in raws], qraw=qraw, scan=synth.raw_to_xyzi(qraw),
                init=(poses[-1] + PERTURB).astype(np.float32), sha256=h.hexdigest()[:16])


def single_config(name, inst, n_ds, m_ds):
    """the `config` object of the headline line — the SAME dict in both arms"""
    _, _, ls, lm = SINGLE_CFGS[name]
    return dict(workload="%s: one synthetic %s scan (%d returns, filters off) vs the local map of %d full-density keyframes at 1 m spacing "
                         "(scan leaf %.2f m, map leaf %.2f m), step = downsampleCurrentScan + kd-tree/grid build + 30 forced LM iterations"
                         % (name, SINGLE_CFGS[name][0], len(inst["scan"]), len(inst["scans"]), ls, lm),
                n_scan=len(inst["scan"]), n_ds=int(n_ds), m_map=int(m_ds), keyframes=len(inst["scans"]), lm_iters=30, input_sha256=inst["sha256"],
                l2="flushed before every timed step (256 MB memset on the GPU; the CPU arm streams a 256 MB array)")


class SingleFrameGpu:
    """the kitti64_single / os1_128_dense instance resident on one GPU"""

    def __init__(self, name, inst, device):
        import torch
        import liorf_b200
        self.torch, self.inst, self.name = torch, inst, name
        _, n_scan, ls, lm = SINGLE_CFGS[name]
        self.ctx = ctx = liorf_b200.Context(device=device, N_SCAN=n_scan, downsampleRate=1, point_filter_num=1, mappingSurfLeafSize=ls, surroundingKeyframeMapLeafSize=lm)
        ctx.reserve(1 << 18, 4 << 20, 4 << 20, 0)
        for k, (sc, p) in enumerate(zip(inst["scans"], inst["poses"])):       # surfCloudKeyFrames[k] = the scan at pose k voxelised by the library itself
            ctx.setCurrentScan(sc)
            ctx.downsampleCurrentScan(want_output=False)
            ctx.addKeyframe(p.astype(np.float32), 0.1 * k)
        self.ids = list(range(len(inst["scans"])))
        self.m_ds = ctx.extractSurroundingKeyFrames(self.ids)
        self.dev = torch.device(f"cuda:{device}")
        self.ext = torch.cuda.ExternalStream(ctx.stream(), device=self.dev)
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=self.dev)
        self.set_scan(inst["scan"])

    def set_scan(self, xyzi):
        t = self.torch
        self.n = len(xyzi)
        self.pin = t.from_numpy(np.ascontiguousarray(xyzi, np.float32)).pin_memory()
        self.d_scan = self.pin.to(self.dev)
        t.cuda.synchronize()
        self.ctx.setCurrentScanDev(self.d_scan.data_ptr(), self.n)

    def run(self, steps, warmup, mode="dev", force_all=True, sections=None):
        """returns per-step device ms (e0→e1 around grid build + downsample + solver [+ H2D]), per-step wall ms (incl. the pose read-back), last pose"""
        import ctypes as C
        t, ctx = self.torch, self.ctx
        dev_ms, wall_ms = [], []
        pose = None
        for i in range(warmup + steps):
            if i == warmup:
                ctx.enableTiming(True, sections=sections)
                self.launches0 = ctx.launchCount()
            with t.cuda.stream(self.ext):
                self.flush.zero_()                                  # L2 flush (256 MB > 126 MB L2), untimed
            ctx.sync()
            e0 = t.cuda.Event(enable_timing=True); e1 = t.cuda.Event(enable_timing=True)
            w0 = time.perf_counter()
            with t.cuda.stream(self.ext):
                e0.record()
            if mode == "e2e":                                       # laserCloudSurfLast from HOST memory: H2D copy of the scan inside the region
                rc = ctx.lib.liorf_set_current_scan(ctx.h, C.c_void_p(self.pin.data_ptr()), C.c_int(self.n))
                assert rc == 0
            ctx.kdtreeSetInputCloud()
            ctx.downsampleCurrentScan(want_output=False)
            ctx.scan2MapOptimizationAsync(self.inst["init"], 30, force_all)
            with t.cuda.stream(self.ext):
                e1.record()
            pose = ctx.getPose()                                    # D2H of the pose (+ counts), the step's only read-back
            w1 = time.perf_counter()
            if i >= warmup:
                dev_ms.append(e0.elapsed_time(e1)); wall_ms.append((w1 - w0) * 1e3)
        self.launches = ctx.launchCount() - self.launches0
        self.timing = ctx.getTiming()
        ctx.enableTiming(False)
        self.counts = ctx.lastCounts()
        return np.array(dev_ms), np.array(wall_ms), pose

    def map_build_ms(self, reps=5):
        """extractSurroundingKeyFrames' device work (transform + concat + VoxelGrid + grid) with the cache invalidated, CUDA events on the map stream's sections"""
        ctx = self.ctx
        kf0 = ctx.getKeyframe(0)[1]
        ctx.enableTiming(True, sections=["map_build", "grid_build"])
        for _ in range(reps):
            ctx.updateKeyframePose(0, kf0)
            ctx.extractSurroundingKeyFrames(self.ids, want_count=False)
            ctx.sync()
        tm = ctx.getTiming(); ctx.enableTiming(False)
        return (tm["map_build"][0] + tm["grid_build"][0]) / max(tm["map_build"][1], 1)

    def close(self):
        self.ctx.close()


def stats_ms(a):
    return dict(mean=float(np.mean(a)), median=float(np.median(a)), p95=float(np.percentile(a, 95)), min=float(np.min(a)), n=int(len(a)))


def traffic_entry(key):
    """profiles/traffic.json: per-launch DRAM / L2 bytes of a kernel on a named workload, from the committed `ncu --set full` capture"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(key)
    except Exception:
        return None


# ----------------------------------------------------------------------------------------------------------------
# config 4: Livox deskew + registration (one frame: projectPointCloud with the 200 Hz IMU table → downsample → grid → solve)
# ----------------------------------------------------------------------------------------------------------------
LIVOX_FILT = dict(lidarMinRange=1.0, lidarMaxRange=1000.0, N_SCAN=6, downsampleRate=1, point_filter_num=3)   # config/lio_sam_livox.yaml:27-32


def livox_inputs(nkf=50):
    """config 4 inputs (seeded): keyframe scans along +x, the query scan swept at 1 rad/s with its 200 Hz IMU table, the start pose"""
    from tools import synth
    poses = [np.array([0, 0, 0, 1.0 * k, 0, 0], np.float64) for k in range(nkf)]
    kf_scans = [synth.raw_to_xyzi(synth.scan(synth.LIVOX, p, seed=synth.SEED0 + 3000 + k)) for k, p in enumerate(poses)]
    omega = (0.02, -0.03, 1.0)                                             # 1 rad/s yaw sweep: the deskew is not a no-op
    qp = poses[-1]
    raw = synth.scan(synth.LIVOX, qp, omega=omega, seed=synth.SEED0 + 3500)
    t0 = 20.0
    it, rot, ptr = synth.imu_table(t0, t0 + float(raw["time"][-1]), omega, rate_hz=200.0, gyro_noise=1e-3, seed=4)
    return dict(poses=poses, kf_scans=kf_scans, raw=raw, t0=t0, imu_time=it, imu_rot=rot, imu_ptr=ptr, init=(qp + 0.5 * PERTURB).astype(np.float32))


def livox_cpu(L, reps, gpu=None):
    """the CPU path on the same config-4 step (projectPointCloud + downsampleCurrentScan + kd-tree build + scan2MapOptimization with the
    reference's convergence break), all host threads; `gpu` = (pose, counts) of the GPU run of the same inputs → differences reported"""
    if ORACLE_DIR not in sys.path:
        sys.path.insert(0, ORACLE_DIR)
    import pyoracle as o
    all_cores = os.cpu_count() or 1
    o.set_num_threads(all_cores)
    kfs = [o.voxel_grid(s, 0.15)[0] for s in L["kf_scans"]]
    mp, _, _ = o.voxel_grid(np.concatenate([o.transform_cloud(c, p.astype(np.float32)) for c, p in zip(kfs, L["poses"])]), 0.3)
    use_ref = o.ref() is not None
    node = None
    try:                                                            # the reference's own mapOptimization member functions where oracle/_ref holds them
        if o.RefMapOpt.available() and os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libliorf_ref_mapopt_omp.so")):
            node = o.RefMapOpt(openmp=True, numberOfCores=all_cores, N_SCAN=6, Horizon_SCAN=4000, mappingSurfLeafSize=0.15, surroundingKeyframeMapLeafSize=0.3)
            node.set_map(mp)
    except Exception:
        node = None
    ts, split = [], np.zeros(3)
    n_ds = iters = 0; tf = None
    for r in range(reps + 1):
        a = time.perf_counter()
        cloud, kept = o.project_point_cloud(L["raw"], LIVOX_FILT, L["t0"], L["imu_time"], L["imu_rot"], L["imu_ptr"], True)
        b = time.perf_counter()
        if node is not None:
            tf, iters, tm = node.bench_step(cloud, L["init"], 30, False)        # downsampleCurrentScan + kd-tree + the loop with the reference's convergence break
            d = time.perf_counter(); c = b + tm["downsample"] * 1e-3
            n_ds = node.state()["n_ds"]
        else:
            ds, _, _ = o.voxel_grid(cloud, 0.15)
            c = time.perf_counter()
            res = o.scan2map(ds, mp, L["init"], 30, False, None, use_ref_kdtree=use_ref)
            d = time.perf_counter()
            tf, iters, n_ds = res["tf"], int(res["iters"]), len(ds)
        if r > 0:
            ts.append((d - a) * 1e3); split += np.array([b - a, c - b, d - c]) * 1e3
    if node is not None:
        node.close()
    out = dict(value=float(np.median(ts)), unit="ms/frame", cores=all_cores, kind="reference" if node is not None else "port",
               sample="%d repetitions of the same livox_deskew step (after 1 warm-up): deskew = oracle port of projectPointCloud / deskewPoint (serial, as the reference), then %s; OpenMP %d threads"
                      % (reps, "the reference's own downsampleCurrentScan + scan2MapOptimization loop (src/mapOptmization.cpp compiled unchanged, -O3 + OpenMP; stand-ins behind its third-party calls)"
                         if node is not None else "the oracle's VoxelGrid + scan2MapOptimization (early exit), kd-tree = the reference's vendored nanoflann%s" % ("" if use_ref else " (oracle/_ref missing: brute-force kNN)"), all_cores),
               ms=stats_ms(ts), split_ms=dict(deskew=split[0] / reps, downsample=split[1] / reps, scan2map=split[2] / reps),
               n_kept=int(len(kept)), n_ds=int(n_ds), m_map=int(len(mp)), lm_iters=int(iters))
    if gpu is not None:
        pose, cnt, m = gpu
        out["gpu_vs_cpu"] = dict(final_pose_max_abs_diff=float(np.max(np.abs(np.asarray(pose, np.float64) - np.asarray(tf, np.float64)))),
                                 same_n_kept=bool(cnt["n_scan"] == len(kept)), same_n_ds=bool(cnt["n_ds"] == n_ds), same_m_map=bool(m == len(mp)),
                                 same_lm_iters=bool(cnt["iters"] == int(iters)))
    return out


def livox_row(device, steps, warmup, cpu_reps=5):
    import ctypes as C
    import torch
    import liorf_b200
    L = livox_inputs()
    nkf = len(L["poses"])
    ctx = liorf_b200.Context(device=device, N_SCAN=6, downsampleRate=1, point_filter_num=3, mappingSurfLeafSize=0.15, surroundingKeyframeMapLeafSize=0.3)
    ctx.reserve(1 << 16, 1 << 20, 1 << 20, 0)
    for k in range(nkf):
        ctx.setCurrentScan(L["kf_scans"][k])
        ctx.downsampleCurrentScan(want_output=False)
        ctx.addKeyframe(L["poses"][k].astype(np.float32), 0.1 * k)
    m = ctx.extractSurroundingKeyFrames(list(range(nkf)))
    raw, t0, it, rot, ptr, init = L["raw"], L["t0"], L["imu_time"], L["imu_rot"], L["imu_ptr"], L["init"]
    dev = torch.device(f"cuda:{device}")
    ext = torch.cuda.ExternalStream(ctx.stream(), device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    d_raw = torch.from_numpy(raw.view(np.uint8).reshape(-1).copy()).to(dev)
    ms = []
    for i in range(warmup + steps):
        if i == warmup:
            ctx.enableTiming(True)
        with torch.cuda.stream(ext):
            flush.zero_()
        ctx.sync()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ext):
            e0.record()
        ctx.projectPointCloudDev(d_raw.data_ptr(), len(raw), t0, it, rot, ptr, True)
        ctx.kdtreeSetInputCloud()
        ctx.downsampleCurrentScan(want_output=False)
        ctx.scan2MapOptimizationAsync(init, 30, False)
        with torch.cuda.stream(ext):
            e1.record()
        pose = ctx.getPose()
        if i >= warmup:
            ms.append(e0.elapsed_time(e1))
    tm = ctx.getTiming(); cnt = ctx.lastCounts()
    ctx.close()
    try:                                                                    # the CPU figure must never take the GPU row down
        cpu = livox_cpu(L, cpu_reps, (pose, cnt, m)) if cpu_reps > 0 else None
    except Exception as e:
        cpu = dict(error=repr(e))
    n_kept = cnt["n_scan"]
    dk_ms = tm["deskew"][0] / max(tm["deskew"][1], 1)
    alg = 24 * len(raw) + 16 * n_kept
    return dict(workload="livox_deskew: %d-point rosette scan, 200 Hz 6-axis IMU table (%d rows), point_filter_num 3, leaves 0.15 / 0.3, %d-keyframe map; "
                         "step = projectPointCloud (deskew) + downsample + grid build + scan2MapOptimization (early exit)" % (len(raw), ptr + 1, nkf),
                n_kept=n_kept, n_ds=cnt["n_ds"], m_map=m, lm_iters=cnt["iters"], frame_ms=stats_ms(ms),
                kernel_ms={k: v[0] / max(v[1], 1) for k, v in tm.items() if v[1]},
                deskew_roofline=dict(kernel="k_first_kept + scan + k_deskew_points", bound="hbm", achieved=alg / (dk_ms * 1e-3) / 1e9 if dk_ms > 0 else None, unit="GB/s",
                                     algorithmic_bytes_per_launch=alg, avg_launch_ms=dk_ms),
                cpu_baseline=cpu)


# ----------------------------------------------------------------------------------------------------------------
# config 2 (extra at N = 1): a drive through liorf_process_frame, device-resident and end to end
# ----------------------------------------------------------------------------------------------------------------
def sequence_extra(local_rank, rank, P, W, K):
    import torch
    dev = torch.device(f"cuda:{local_rank}")
    n_frames = P + W + K + 1
    seq = Sequence(n_frames, rank)
    for i in range(n_frames):
        seq.frame(i)
    out = {}
    for mode in ("e2e", "dev"):
        pipe = GpuPipeline(seq, local_rank)
        pipe.stage(range(n_frames))
        for i in range(P):
            pipe.step(i, "dev")
        for i in range(P, P + W):
            pipe.step(i, mode)
        pipe.stats = {k: 0 for k in pipe.stats}
        pipe.ctx.enableTiming(True)
        ext = torch.cuda.ExternalStream(pipe.ctx.stream(), device=dev)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        with torch.cuda.stream(ext):
            e0.record()
        for i in range(P + W, P + W + K):
            pipe.step(i, mode)
        with torch.cuda.stream(ext):
            e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - w0
        tm = pipe.ctx.getTiming()
        st = pipe.stats
        out[mode] = dict(ms_per_frame=e0.elapsed_time(e1) / K, wall_ms_per_frame=wall * 1e3 / K, frames=K, keyframes_added=st["keyframes"],
                         avg_n_ds=st["n_ds"] / max(st["frames"], 1), avg_m_ds=st["m_ds"] / max(st["frames"], 1), avg_lm_iters=st["iters"] / max(st["frames"], 1),
                         kernel_ms_per_frame={k: v[0] / K for k, v in tm.items() if v[1]})
        pipe.ctx.close()
    try:                                                            # the CPU figure must never take the GPU row down
        cpu = sequence_cpu_reference(seq, P + W, min(K, 20))
    except Exception as e:
        cpu = dict(error=repr(e))
    return dict(workload="kitti05_seq: synthetic 64-beam drive, %d-frame window after a %d-frame pre-roll, yaml filters, early-exit LM, one liorf_process_frame call per frame "
                         "with look-ahead (next frame's H2D + deskew + downsample overlap this frame's solve); every section timed (costs ~10 us/frame of host time)" % (K, P),
                resident=out["dev"], e2e=out["e2e"], h2d_bytes_per_frame=int(np.mean([seq.raw[i].nbytes for i in range(P + W, P + W + K)])), d2h_bytes_per_frame=64,
                cpu_baseline=cpu)


def sequence_cpu_reference(seq, preroll, frames, threads=None):
    """config 2 on the host cores: the same frames, one liorf::cloud_info per scan through the reference's OWN mapOptimization node (src/mapOptmization.cpp compiled unchanged
    with -O3 + OpenMP, oracle/_ref/libliorf_ref_mapopt_omp.so): updateInitialGuess, extractSurroundingKeyFrames, downsampleCurrentScan, scan2MapOptimization (early exit),
    saveKeyFramesAndFactor incl. the ScanContext descriptor.  The deskew in front of it is the oracle port of projectPointCloud (the reference's ImageProjection node gives the same
    bits, tests/test_oracle_vs_reference_nodes.py).  `preroll` frames build the keyframe map untimed, `frames` frames are timed."""
    if ORACLE_DIR not in sys.path:
        sys.path.insert(0, ORACLE_DIR)
    import pyoracle as o
    threads = threads or (os.cpu_count() or 1)
    if not (o.RefMapOpt.available() and os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libliorf_ref_mapopt_omp.so"))):
        return dict(unavailable="oracle/_ref/libliorf_ref_mapopt_omp.so is not on this box")
    R = o.RefMapOpt(openmp=True, numberOfCores=threads, useImuHeadingInitialization=1)
    prev, t_dk, t_node, kf0 = None, [], [], 0
    for i in range(preroll + frames):
        raw, (t0, it, rot, ptr) = seq.frame(i)
        a = time.perf_counter()
        cloud, _ = o.project_point_cloud(raw, seq.filters, t0, it, rot, ptr, True)
        b = time.perf_counter()
        g = np.asarray(seq.initial_guess(i, prev), np.float32)
        R.set_transform(g)                                          # transformTobeMapped = this frame's guess (no odometry / IMU increment on top)
        R.cloud_info(t0, cloud, 0, 0, g[:3], None)
        c = time.perf_counter()
        st = R.state(); prev = st["tf"].copy()
        if i == preroll - 1:
            kf0 = st["keyframes"]
        if i >= preroll:
            t_dk.append((b - a) * 1e3); t_node.append((c - b) * 1e3)
    R.close()
    tot = np.array(t_dk) + np.array(t_node)
    return dict(value=float(np.mean(tot)), unit="ms/frame", cores=threads, kind="reference",
                sample="%d frames after a %d-frame pre-roll (the frames the GPU window starts with): deskew = oracle port of projectPointCloud (1 thread, as the reference), then the "
                       "reference's own mapOptimization node (laserCloudInfoHandler, -O3 + OpenMP %d threads; stand-ins behind its PCL / Eigen / OpenCV calls)" % (frames, preroll, threads),
                ms=stats_ms(list(tot)), split_ms=dict(deskew=float(np.mean(t_dk)), map_optimization_node=float(np.mean(t_node))), keyframes_added=int(st["keyframes"] - kf0),
                final_pose=[float(v) for v in prev])


# ----------------------------------------------------------------------------------------------------------------
# config 5: ScanContext search, database sharded over the ranks
# ----------------------------------------------------------------------------------------------------------------
class ScBench:
    def __init__(self, device, rank, world, K, q_max, dist, lanes):
        import torch
        import liorf_b200
        from liorf_b200.sc_sharded import PeerShardedSearch
        from tools import synth
        self.torch, self.synth, self.rank, self.world, self.K, self.dist = torch, synth, rank, world, K, dist
        self.bounds = [K * g // world for g in range(world + 1)]
        lo, hi = self.bounds[rank], self.bounds[rank + 1]
        self.dev = torch.device(f"cuda:{device}")

        def load(c, a, b):
            c.reserve(1024, 1024, 0, max(b - a, 1))
            for s in range(a, b, 10000):
                c.scAddDescriptors(synth.sc_descriptors(min(10000, b - s), first=s))
        self.owner = liorf_b200.Context(device=device)
        load(self.owner, lo, hi)
        peer = PeerShardedSearch(self.owner, rank, world, self.bounds, q_max, torch)
        peer.connect_processes(dist)                               # maps the windows and replicates the ring keys (NVLink pushes by the library's kernels)
        self.owner.sync()
        self.lanes = [(self.owner, peer)]
        for _ in range(max(1, lanes) - 1):                         # further batches in flight: contexts that borrow the database and the replicated index
            c2 = liorf_b200.Context(device=device)
            c2.scBorrowDatabase(self.owner)
            p2 = PeerShardedSearch(c2, rank, world, self.bounds, q_max, torch, k_total_max=0)
            p2.connect_processes(dist)
            self.lanes.append((c2, p2))
        # the unsharded search on rank 0 (full copy of the database, same number of lanes): the bit-equality oracle of the run and the one-GPU figure
        self.ref_lanes = []
        if world > 1 and rank == 0:
            full = liorf_b200.Context(device=device)
            load(full, 0, K)
            self.ref_lanes = [(full, PeerShardedSearch(full, 0, 1, [0, K], q_max, torch))]
            for _ in range(max(1, lanes) - 1):
                c2 = liorf_b200.Context(device=device); c2.scBorrowDatabase(full)
                self.ref_lanes.append((c2, PeerShardedSearch(c2, 0, 1, [0, K], q_max, torch)))
            for _, s in self.ref_lanes:
                s.connect_local([s])
        n_src = min(K, 2000)
        self.src_rows = (np.arange(n_src, dtype=np.int64) * K) // n_src     # loop sources spread over ALL rows (and so over all ranks)
        self.sample = np.concatenate([synth.sc_descriptors(1, first=int(i)) for i in self.src_rows])

    def queries(self, Q):
        qd, src, shift = self.synth.sc_queries(self.sample, Q)
        return qd, np.where(src >= 0, self.src_rows[np.maximum(src, 0)], -1), shift

    def _timed(self, lanes, d_q, n_batches, warm, barrier):
        t = self.torch
        streams = [s.stream for _, s in lanes]
        turn = 0
        for _ in range(warm):
            lanes[turn % len(lanes)][1].query(d_q); turn += 1
        for c, _ in lanes:
            c.sync()
        lanes[0][1].wait_stats()
        if barrier:
            self.dist.barrier()
        t.cuda.synchronize()
        e0 = t.cuda.Event(enable_timing=True); e1 = t.cuda.Event(enable_timing=True)
        with t.cuda.stream(streams[0]):
            e0.record()
        for st in streams[1:]:
            st.wait_event(e0)                                      # every lane starts inside the timed region
        out = None
        for b in range(n_batches):
            out = lanes[b % len(lanes)][1].query(d_q)
        for st in streams[1:]:
            streams[0].wait_stream(st)                             # ... and ends inside it
        with t.cuda.stream(streams[0]):
            e1.record()
        for c, _ in lanes:
            c.sync()
        t.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        waits, _ = lanes[0][1].wait_stats()
        return ms, waits, out

    def sync_lanes(self):
        try:
            for c, _ in self.lanes:
                c.sync()
        except Exception:
            if os.environ.get("LIORF_BENCH_VERBOSE"):
                time.sleep(15)                                     # let the peers run into their own bounds, then show every lane's view
                for k, (c, s) in enumerate(self.lanes):
                    print("[bench rank %d] lane %d state %s" % (self.rank, k, s.debug_state()), file=sys.stderr, flush=True)
            raise

    def note(self, msg):
        if os.environ.get("LIORF_BENCH_VERBOSE"):
            print("[bench rank %d] %s" % (self.rank, msg), file=sys.stderr, flush=True)

    def run(self, Q, steps, warmup):
        """steps batches of Q queries, round-robin over the lanes; device-timed, max over ranks.  Returns a result dict (rank 0: + checks)."""
        t, dist, world = self.torch, self.dist, self.world
        qd, src, shift = self.queries(Q)
        pin = t.from_numpy(qd).pin_memory()
        d_q = pin.to(self.dev)
        t.cuda.synchronize()
        if world > 1:
            dist.barrier()
        warm = max(warmup, 3) * len(self.lanes)                    # per lane: the third identical request captures the batch as a CUDA graph
        self.note("run Q=%d steps=%d: timed region" % (Q, steps))
        if os.environ.get("LIORF_BENCH_VERBOSE"):
            import faulthandler
            faulthandler.dump_traceback_later(8, exit=False)
        ms, waits, out = self._timed(self.lanes, d_q, steps, warm, world > 1)
        if world > 1:
            tt = t.tensor([ms], device=self.dev); dist.all_reduce(tt, op=dist.ReduceOp.MAX); ms = float(tt.item())
        res = dict(K=self.K, Q=Q, shards=world, batches_in_flight=len(self.lanes), steps=steps, ms_per_batch=ms / steps, queries_per_s=Q * steps / (ms * 1e-3),
                   peer_wait_us_per_batch={k: v / 1e3 / max(steps / len(self.lanes), 1) for k, v in waits.items()})
        loop, sh, dd, cand = [x.cpu().numpy() for x in out]
        ok = (loop == src) & (src >= 0)
        res.update(planted_loops_found=int(ok.sum()), planted=int((src >= 0).sum()), shifts_correct=int((sh[ok] == shift[ok]).sum()))
        self.note("timed region done")
        if os.environ.get("LIORF_BENCH_VERBOSE"):
            import faulthandler
            faulthandler.cancel_dump_traceback_later()
        skip = os.environ.get("LIORF_BENCH_SKIP", "")              # debugging aid: comma list of ref, stage
        # every rank's answers against the unsharded search of the same queries (rank 0, full database): the real cudaIpc path, bit for bit
        if world > 1 and "ref" not in skip:
            ref = None
            if self.rank == 0:
                rms, _, rout = self._timed(self.ref_lanes, d_q, max(2, steps // 2), 3 * len(self.ref_lanes), False)
                res["unsharded_same_run"] = dict(queries_per_s=Q * max(2, steps // 2) / (rms * 1e-3), ms_per_batch=rms / max(2, steps // 2), batches_in_flight=len(self.ref_lanes),
                                                 note="rank 0 alone, full copy of the database, same batches in flight, while the other ranks idle")
                ref = [x.clone() for x in rout]
            else:
                ref = [t.empty_like(x) for x in out]
            for x in ref:
                dist.broadcast(x, src=0)
            same = all(bool(t.equal(a.view(t.int64) if a.dtype == t.float64 else a, b.view(t.int64) if b.dtype == t.float64 else b)) for a, b in zip(out, ref))
            flag = t.tensor([1 if same else 0], device=self.dev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            res["bit_equal_unsharded"] = bool(flag.item() == 1)
            self.note("bit-equality check done: %s" % res["bit_equal_unsharded"])
        return res

    def run_e2e(self, Q, n_batches):
        """the same batches through the HOST-buffer entry of the C ABI (liorf_sc_shard_query_async): every batch copies its query descriptors from
        pinned host memory to the device and its answers back, inside the timed region; round-robin over the lanes; device-timed, max over ranks"""
        t, dist, world = self.torch, self.dist, self.world
        qd, src, shift = self.queries(Q)
        pin = t.from_numpy(qd).pin_memory()
        host_out = [(t.empty(Q, dtype=t.int32).pin_memory(), t.empty(Q, dtype=t.int32).pin_memory(), t.empty(Q, dtype=t.float64).pin_memory(),
                     t.empty((Q, 3), dtype=t.int32).pin_memory()) for _ in self.lanes]
        streams = [s.stream for _, s in self.lanes]
        for k, (c, s) in enumerate(self.lanes):                   # warm-up: sizes the library's staging buffer of every lane
            s.query_host(pin, host_out[k])
        self.sync_lanes()
        if world > 1:
            dist.barrier()
        t.cuda.synchronize()
        e0 = t.cuda.Event(enable_timing=True); e1 = t.cuda.Event(enable_timing=True)
        with t.cuda.stream(streams[0]):
            e0.record()
        for st in streams[1:]:
            st.wait_event(e0)
        for b in range(n_batches):
            k = b % len(self.lanes)
            self.lanes[k][1].query_host(pin, host_out[k])
        for st in streams[1:]:
            streams[0].wait_stream(st)
        with t.cuda.stream(streams[0]):
            e1.record()
        self.sync_lanes()
        t.cuda.synchronize()
        ems = e0.elapsed_time(e1)
        if world > 1:
            tt = t.tensor([ems], device=self.dev); dist.all_reduce(tt, op=dist.ReduceOp.MAX); ems = float(tt.item())
        loop = host_out[(n_batches - 1) % len(self.lanes)][0].numpy()
        ok = (loop == src) & (src >= 0)
        self.note("e2e done")
        return dict(value=Q * n_batches / (ems * 1e-3), unit="queries/s", h2d_bytes_per_step=int(qd.nbytes), d2h_bytes_per_step=int(Q * (4 + 4 + 8 + 12)), batches=n_batches,
                    planted_loops_found=int(ok.sum()), planted=int((src >= 0).sum()),
                    note="liorf_sc_shard_query_async: every rank copies the batch's query descriptors (9 600 B each) from pinned host memory and reads all answers back — PCIe-bound")

    def stage_times(self, Q, reps=6):
        """live CUDA-event sections of lane 0 (plain launches): ring-key stage, the tcgen05 GEMM inside it"""
        t = self.torch
        ctx, peer = self.lanes[0]
        qd, _, _ = self.queries(Q)
        d_q = t.from_numpy(qd).to(self.dev)
        t.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        ctx.enableTiming(True)
        for _ in range(reps):
            peer.query(d_q)
        ctx.sync()
        tm = ctx.getTiming(); ctx.enableTiming(False)
        st = ctx.scTensorStats()
        if self.world > 1:
            self.dist.barrier()
        return tm, st

    def close(self):
        for c, _ in self.lanes[1:] + self.ref_lanes[1:]:
            c.close()
        for c, _ in self.ref_lanes[:1]:
            c.close()
        self.owner.close()


def sc_rooflines(scb, Q, peaks):
    """tensor roofline of the ring-key GEMM and HBM roofline of stage 2, both timed live (CUDA events) on this rank's share of a batch"""
    import ctypes as C
    t = scb.torch
    if "stage" in os.environ.get("LIORF_BENCH_SKIP", ""):
        return dict(roofline=None)
    tm, st = scb.stage_times(Q)
    Qs = Q // scb.world if scb.world > 1 else Q
    gemm_ms = tm["sc_gemm"][0] / max(tm["sc_gemm"][1], 1)
    kpad, qpad = (scb.K + 127) // 128 * 128, (Qs + 255) // 256 * 256
    tflops = 2.0 * 64 * kpad * qpad / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
    out = dict(ringkey_stage_ms=tm["sc_search"][0] / max(tm["sc_search"][1], 1), candidates_per_query=st["candidates"] / max(Qs, 1) * 32, overflow_queries=st["overflow"],
               roofline=dict(kernel="k_sc_tensor", bound="tensor", achieved=tflops, peak=peaks.get("bf16_tflops"), unit="TFLOP/s",
                             frac=(tflops / peaks["bf16_tflops"]) if tflops and peaks.get("bf16_tflops") else None,
                             traffic=(traffic_entry("sc_tensor.k%dk_q%d" % (scb.K // 1000, Qs)) or {}).get("dram_bytes_per_launch"), avg_launch_ms=gemm_ms,
                             queries_per_launch=Qs, keys=scb.K,
                             note="executed tensor-core flops: 2 x 64 (split-bf16 contraction) x Kpad x Qpad per launch; the distance itself is 3 x 20 flops per pair"))
    if scb.world == 1:
        ctx, peer = scb.lanes[0]
        qd, _, _ = scb.queries(Q)
        with t.cuda.stream(peer.stream):
            d_q = t.from_numpy(qd).to(scb.dev)
        loop, sh, dd, cand = peer.query(d_q)
        ctx.sync()
        pd_ = t.empty((Q, 3), dtype=t.float64, device=scb.dev); ps_ = t.empty((Q, 3), dtype=t.int32, device=scb.dev)
        vp = lambda x: C.c_void_p(x.data_ptr())
        call = lambda: ctx.lib.liorf_sc_distance_batch_dev(ctx.h, vp(d_q), vp(cand), Q, 0, vp(pd_), vp(ps_))
        for _ in range(2):
            call()
        ctx.sync()
        s0 = t.cuda.Event(enable_timing=True); s1 = t.cuda.Event(enable_timing=True)
        with t.cuda.stream(peer.stream):
            s0.record()
        for _ in range(5):
            call()
        with t.cuda.stream(peer.stream):
            s1.record()
        ctx.sync()
        s2_ms = s0.elapsed_time(s1) / 5
        alg = 3 * Q * (9600 + 9600 / 3 + 2 * 480)                # candidate descriptor + a third of the query's + the candidate's sector key and column norms
        tr = traffic_entry("sc_distance_bulk.k%dk_q%d" % (scb.K // 1000, Q))
        out["stage2_roofline"] = dict(kernel="k_sc_distance_bulk", bound="hbm", achieved=alg / (s2_ms * 1e-3) / 1e9, peak=peaks["hbm_gbs"], unit="GB/s",
                                      frac=alg / (s2_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], traffic=(tr or {}).get("dram_bytes_per_launch"), algorithmic_bytes_per_launch=alg,
                                      avg_launch_ms=s2_ms, pairs=3 * Q,
                                      note="fp64 sums in the reference's sequential order; one warp per pair, adjacent columns / shifts blocked in registers, 16 pairs in flight per SM")
    return out


# ----------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sc-k", type=int, default=100000)
    ap.add_argument("--sc-q", type=int, default=32768, help="queries per batch of the ScanContext workload (the N > 1 headline)")
    ap.add_argument("--sc-q-sweep", default="4096,131072", help="further batch sizes reported as extras (empty = none)")
    ap.add_argument("--sc-lanes", type=int, default=6, help="ScanContext query batches in flight per GPU — the SAME at every N")
    ap.add_argument("--no-sc", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="N = 1: headline only (no rows / sequence / sc extras)")
    ap.add_argument("--seq-frames", type=int, default=60, help="timed frames of the kitti05_seq extra (0 = skip)")
    ap.add_argument("--seq-preroll", type=int, default=100)
    ap.add_argument("--cpu-reps", type=int, default=20, help="bounded CPU-baseline sample: repetitions per thread setting (BASELINE.md §3: median + p95 over >= 20)")
    args = ap.parse_args()
    rank, local_rank, world = dist_env()
    W = max(args.warmup, 3)
    K = args.steps

    if args.impl == "reference":
        if rank != 0:
            return 0
        return run_reference(args, W, K, world)

    import torch
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    torch.cuda.set_device(local_rank)
    peaks, peak_src = load_peaks()
    rc = bench_sc_headline(args, rank, local_rank, world, W, K, dist, peaks, peak_src) if world > 1 else bench_single_headline(args, local_rank, W, K, peaks, peak_src)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return rc


def bench_single_headline(args, device, W, K, peaks, peak_src):
    import torch
    name = "kitti64_single"
    inst = make_single_inputs(name)
    sf = SingleFrameGpu(name, inst, device)
    sampler = ClockSampler(device)
    e2e_dev, e2e_wall, _ = sf.run(K, W, "e2e", sections=["scan2map"])
    e2e_launch = sf.launches
    sampler.start()
    dev_ms, wall_ms, pose = sf.run(K, W, "dev", sections=["scan2map"])
    clocks = sampler.stop()
    launches = sf.launches
    s2m = sf.timing["scan2map"]
    s2m_ms = s2m[0] / max(s2m[1], 1)
    cnt = sf.counts
    n_ds, m_ds = cnt["n_ds"], cnt["m_ds"]
    # per-kernel split of the step (every section timed: a second, shorter pass)
    _, _, _ = sf.run(min(K, 10), 1, "dev")
    split = {k: v[0] / max(v[1], 1) for k, v in sf.timing.items() if v[1]}
    alg = 96.0 * n_ds * 30                                       # 16 B query point + 5 x 16 B neighbours per query and iteration
    tr = traffic_entry("scan2map.kitti64_single") or {}
    roofline = dict(kernel="k_scan2map_persistent", bound="hbm", achieved=alg / (s2m_ms * 1e-3) / 1e9, peak=peaks["hbm_gbs"], unit="GB/s",
                    frac=alg / (s2m_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], traffic=tr.get("dram_bytes_per_launch"), algorithmic_bytes_per_launch=alg,
                    peak_source=peak_src, avg_launch_ms=s2m_ms, share_of_step=s2m_ms / float(np.mean(dev_ms)),
                    l2_bytes_per_launch=tr.get("l2_bytes_per_launch"),
                    l2_gbs=(tr["l2_bytes_per_launch"] / (s2m_ms * 1e-3) / 1e9) if tr.get("l2_bytes_per_launch") else None,
                    note="30 dependent {gather -> grid-wide reduce -> 6x6 solve} rounds on an L2-resident working set: latency-bound by construction (SURVEY §8d note); "
                         "the HBM fraction is reported as the contract asks, the L2 figure is what the gathers actually move (ncu lts__t_bytes of the committed capture)")
    line = dict(metric="scan2map_ms_per_frame_64beam", value=float(np.mean(dev_ms)), unit="ms/frame", n_gpus=1, steps=K, warmup=W, ms_per_step=float(np.mean(dev_ms)),
                higher_is_better=False, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", config=single_config(name, inst, n_ds, m_ds),
                solve_ms=stats_ms(dev_ms), wall_ms_per_step=stats_ms(wall_ms), target_ms=1.0,
                e2e=dict(value=float(np.mean(e2e_wall)), unit="ms/frame", h2d_bytes_per_step=int(sf.n * 16 + 4), d2h_bytes_per_step=6 * 4 + 2 * 16 * 4 + 4 * 4,
                         device_ms=stats_ms(e2e_dev), wall_ms=stats_ms(e2e_wall), gpu_launches=e2e_launch,
                         timing="host wall clock around one step: liorf_set_current_scan (pinned host scan, H2D) + grid build + downsample + solve + liorf_get_pose (D2H)"),
                gpu_launches=launches, kernel_ms_per_step=split, knn_queries_per_s=30.0 * n_ds / (s2m_ms * 1e-3), final_pose=[float(v) for v in pose],
                roofline=roofline, clocks=clocks)
    rows = {}
    if not args.no_extras:
        rows["map_build_ms"] = sf.map_build_ms()
        d2, _, _ = sf.run(min(K, 20), 2, "dev", force_all=False)
        rows["early_exit"] = dict(solve_ms=stats_ms(d2), lm_iters=sf.counts["iters"], note="same step with the reference's convergence break (src/mapOptmization.cpp:1313)")
        qraw = inst["qraw"]
        keep = (qraw["ring"] % 2 == 0) & (np.arange(len(qraw)) % 5 == 0)          # config/kitti.yaml: downsampleRate 2, point_filter_num 5
        sf.set_scan(inst["scan"][keep])
        d3, _, _ = sf.run(min(K, 20), 2, "dev")
        rows["yaml_filters"] = dict(solve_ms=stats_ms(d3), n_scan=int(keep.sum()), n_ds=sf.counts["n_ds"], note="same map, the scan decimated by the yaml filters (1/10), 30 forced iterations")
    # ---- CPU baseline on the same step (bounded sample) ----
    line["cpu_baseline"] = cpu_single(inst, name, args.cpu_reps, sf, check_pose=pose)
    sf.close()
    if not args.no_extras:
        try:
            i3 = make_single_inputs("os1_128_dense")
            s3 = SingleFrameGpu("os1_128_dense", i3, device)
            d, w, p3 = s3.run(min(K, 20), 3, "dev")
            s3m = s3.timing["scan2map"]; s3ms = s3m[0] / max(s3m[1], 1)
            rows["os1_128_dense"] = dict(config=single_config("os1_128_dense", i3, s3.counts["n_ds"], s3.counts["m_ds"]), solve_ms=stats_ms(d),
                                         kernel_ms_per_step={k: v[0] / max(v[1], 1) for k, v in s3.timing.items() if v[1]}, map_build_ms=s3.map_build_ms(3),
                                         roofline=dict(kernel="k_scan2map_persistent", bound="hbm", achieved=96.0 * s3.counts["n_ds"] * 30 / (s3ms * 1e-3) / 1e9, peak=peaks["hbm_gbs"],
                                                       unit="GB/s", frac=96.0 * s3.counts["n_ds"] * 30 / (s3ms * 1e-3) / 1e9 / peaks["hbm_gbs"], avg_launch_ms=s3ms),
                                         cpu_baseline=cpu_single(i3, "os1_128_dense", 2, s3, check_pose=p3))
            s3.close()
        except Exception as e:                                      # an extra must never take the headline down
            rows["os1_128_dense"] = dict(error=repr(e))
        try:
            rows["livox_deskew"] = livox_row(device, min(K, 20), 3)
        except Exception as e:
            rows["livox_deskew"] = dict(error=repr(e))
        line["rows"] = rows
        if args.seq_frames > 0:
            try:
                line["sequence"] = sequence_extra(device, 0, args.seq_preroll, 5, args.seq_frames)
            except Exception as e:
                line["sequence"] = dict(error=repr(e))
        if not args.no_sc:
            try:
                scb = ScBench(device, 0, 1, args.sc_k, max([args.sc_q] + [int(x) for x in args.sc_q_sweep.split(",") if x]), None, args.sc_lanes)
                sc = scb.run(args.sc_q, 12, 3)
                sc.update(sc_rooflines(scb, args.sc_q, peaks))
                sc["sweep"] = {}
                for q in [int(x) for x in args.sc_q_sweep.split(",") if x]:
                    r = scb.run(q, 12 if q <= 32768 else 4, 3)
                    sc["sweep"][str(q)] = {k: r[k] for k in ("Q", "ms_per_batch", "queries_per_s", "batches_in_flight", "planted_loops_found", "planted")}
                sc["e2e"] = scb.run_e2e(args.sc_q, 6)
                try:                                                # config 5's CPU figure (bounded: 3 steps of 256 queries)
                    if ORACLE_DIR not in sys.path:
                        sys.path.insert(0, ORACLE_DIR)
                    import pyoracle as o
                    sc["cpu_baseline"] = sc_cpu_reference(o, args.sc_k, args.sc_q, 256, 3, 1, os.cpu_count() or 1)[0]
                except Exception as e:
                    sc["cpu_baseline"] = dict(error=repr(e))
                sc["metric"] = "sc_queries_per_s_100k"
                line["sc"] = sc
                line["sc_queries_per_s_100k"] = sc["queries_per_s"]
                scb.close()
            except Exception as e:
                line["sc"] = dict(error=repr(e))
    print(json.dumps(line))
    return 0


REF_NODE_SAMPLE = ("the reference's own src/mapOptmization.cpp compiled unchanged with its build flags (-O3, OpenMP; oracle/_ref/libliorf_ref_mapopt_omp.so): "
                   "downsampleCurrentScan + kd-tree setInputCloud + 30 x {surfOptimization, combineOptimizationCoeffs, LMOptimization}; behind its PCL / Eigen / OpenCV "
                   "calls stand the oracle's restatements, the kd-tree is the reference's vendored nanoflann")


class CpuStep:
    """one headline step on the host cores: the reference's own member functions when oracle/_ref holds the compiled node (kind "reference"), else the
    oracle port (kind "port").  step() -> (final pose, {downsample, kdtree_build, surf_optimization, lm_optimization} in ms)"""

    def __init__(self, o, name, mp, threads):
        self.o, self.mp, self.threads = o, mp, threads
        _, n_scan, self.ls, _ = SINGLE_CFGS[name]
        self.node = None
        try:
            if o.RefMapOpt.available() and os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libliorf_ref_mapopt_omp.so")):
                self.node = o.RefMapOpt(openmp=True, numberOfCores=threads, mappingSurfLeafSize=self.ls, N_SCAN=n_scan)
                self.node.set_map(mp)
        except Exception as e:                                      # a CPU baseline must never take the GPU line down: fall back to the port and say so
            print("[bench] reference node unavailable (%r): CPU baseline falls back to the oracle port" % (e,), file=sys.stderr)
            self.node = None
        if self.node is None:
            o.set_num_threads(threads)
        self.use_ref = o.ref() is not None
        self.kind = "reference" if self.node is not None else "port"

    def step(self, scan, init):
        if self.node is not None:
            pose, _, tm = self.node.bench_step(scan, init, 30, True)
            return pose, tm, self.node.state()["n_ds"]
        a = time.perf_counter()
        ds, _, _ = self.o.voxel_grid(scan, self.ls)
        b = time.perf_counter()
        res = self.o.scan2map(ds, self.mp, init, 30, True, None, use_ref_kdtree=self.use_ref)
        t = res.get("timings", (0.0, 0.0, 0.0))
        return res["tf"], dict(downsample=(b - a) * 1e3, kdtree_build=t[0] * 1e3, surf_optimization=t[1] * 1e3, lm_optimization=t[2] * 1e3), len(ds)

    def close(self):
        if self.node is not None:
            self.node.close()

    def sample(self, name, reps, warm):
        if self.node is not None:
            return "%d repetitions of the same %s step (after %d warm-up): %s; OpenMP %d threads (numberOfCores)" % (reps, name, warm, REF_NODE_SAMPLE, self.threads)
        return "%d repetitions of the same %s step (after %d warm-up): oracle restatement, kd-tree = the reference's vendored nanoflann%s; OpenMP %d threads" % (
            reps, name, warm, "" if self.use_ref else " (oracle/_ref missing: brute-force kNN)", self.threads)


def cpu_single(inst, name, reps, sf=None, check_pose=None):
    """the CPU path on the same step: VoxelGrid of the scan + kd-tree build (the reference's vendored nanoflann, leaf 15) + 30 forced
    iterations with OpenMP over the points (src/mapOptmization.cpp:1078), all host threads and 4 threads (numberOfCores, config/kitti.yaml:63)"""
    if ORACLE_DIR not in sys.path:
        sys.path.insert(0, ORACLE_DIR)
    import pyoracle as o
    _, _, ls, lm = SINGLE_CFGS[name]
    if sf is not None:                                              # keyframe clouds as the GPU run stored them (bit-equal to the oracle's VoxelGrid, tests/test_gpu_parity.py)
        kfs = [sf.ctx.getKeyframe(k)[0] for k in range(len(inst["scans"]))]
    else:
        kfs = [o.voxel_grid(s, ls)[0] for s in inst["scans"]]
    mp, _, _ = o.voxel_grid(np.concatenate([o.transform_cloud(c, p.astype(np.float32)) for c, p in zip(kfs, inst["poses"])]), lm)
    out = {}
    all_cores = os.cpu_count() or 1
    pose = None
    kind, sample, n_ds = "port", "", 0
    for threads in sorted({all_cores, min(4, all_cores)}, reverse=True):
        cs = CpuStep(o, name, mp, threads)
        ts, split = [], {}
        for r in range(reps + 1):
            a = time.perf_counter()
            pose, tm, n_ds = cs.step(inst["scan"], inst["init"])
            c = time.perf_counter()
            if r > 0:
                ts.append((c - a) * 1e3)
                for k, v in tm.items():
                    split[k] = split.get(k, 0.0) + v / reps
        out[threads] = dict(ms=stats_ms(ts), split_ms=split)
        if threads == all_cores:
            kind, sample = cs.kind, cs.sample(name, reps, 1)
        cs.close()
    o.set_num_threads(all_cores)
    best = out[all_cores]
    d = dict(value=best["ms"]["median"], unit="ms/frame", cores=all_cores, kind=kind, sample=sample,
             ms=best["ms"], split_ms=best["split_ms"], n_ds=int(n_ds), m_map=len(mp))
    if 4 in out and all_cores != 4:
        d["threads_4"] = out[4]
    if check_pose is not None and pose is not None:
        d["gpu_vs_cpu_final_pose_max_abs_diff"] = float(np.max(np.abs(np.asarray(check_pose, np.float64) - pose.astype(np.float64))))
    return d


def bench_sc_headline(args, rank, local_rank, world, W, K, dist, peaks, peak_src):
    import torch
    sweep = [int(x) for x in args.sc_q_sweep.split(",") if x]
    scb = ScBench(local_rank, rank, world, args.sc_k, max([args.sc_q] + sweep), dist, args.sc_lanes)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    res = scb.run(args.sc_q, K, W)
    clocks = sampler.stop() if rank == 0 else None
    res.update(sc_rooflines(scb, args.sc_q, peaks))
    res["sweep"] = {}
    for q in sweep:
        r = scb.run(q, max(4, K // 2) if q <= 32768 else 4, 3)
        res["sweep"][str(q)] = {k: r.get(k) for k in ("Q", "ms_per_batch", "queries_per_s", "batches_in_flight", "planted_loops_found", "planted", "bit_equal_unsharded",
                                                       "unsharded_same_run", "peer_wait_us_per_batch")}
    ok = res.get("bit_equal_unsharded", False) and all(v.get("bit_equal_unsharded", False) for v in res["sweep"].values())
    res["e2e"] = scb.run_e2e(args.sc_q, max(2, min(K, 8)))
    if os.environ.get("LIORF_BENCH_VERBOSE"):                      # a sharded run AFTER the host-buffer batches must still work
        r2 = scb.run(4096, 8, 3)
        scb.note("post-e2e run: %.1f M queries/s, bit-equal %s" % (r2["queries_per_s"] / 1e6, r2.get("bit_equal_unsharded")))
    scb.close()
    if rank == 0:
        e2e = res.pop("e2e")
        line = dict(metric="sc_queries_per_s_100k", value=res["queries_per_s"], unit="queries/s", n_gpus=world, steps=K, warmup=W, ms_per_step=res["ms_per_batch"],
                    higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32 ring keys (bf16 split filter + exact re-rank), f64 descriptors", data="synthetic",
                    config=sc_config(args),
                    e2e=e2e, gpu_launches=11 * K, sc=res, roofline=res["roofline"], clocks=clocks,
                    note="N > 1 measures the one workload of the path that shards (SURVEY §8e); the N = 1 line's headline is kitti64_single and carries the same search on "
                         "one GPU, same batches in flight, as `sc` / `sc_queries_per_s_100k`; `sc.unsharded_same_run` is that figure measured in THIS run on rank 0")
        u = res.get("unsharded_same_run") or {}
        if u.get("queries_per_s"):                                  # the one-GPU figure of THIS metric, measured in this run (the driver's N = 1 line carries the other headline)
            line["one_gpu_same_metric"] = dict(metric="sc_queries_per_s_100k", value=u["queries_per_s"], unit="queries/s", n_gpus=1,
                                               batches_in_flight=u.get("batches_in_flight"), measured="rank 0 alone, in this run, full copy of the database")
        print(json.dumps(line))
    return 0 if ok else 3


def sc_config(args):
    return dict(workload="sc_100k: ScanContext loop-closure search, %d-keyframe synthetic database (descriptor rows sharded by contiguous ranges over the ranks, 80-byte ring keys "
                         "replicated), one step = one batch of %d replicated queries (50 %% column-shifted noisy copies of database rows, 50 %% fresh)" % (args.sc_k, args.sc_q),
                K=args.sc_k, Q=args.sc_q, l2="inputs larger than L2: a batch reads 315 MB of query descriptors and ~0.9 GB of candidate descriptors")


def sc_cpu_reference(o, Kdb, sc_q, Qc, steps, warm, all_cores):
    """config 5 on the host cores: Qc of a batch's sc_q queries per step against the same Kdb-entry database — every query through the reference's own
    SCManager::detectLoopClosureID when oracle/_ref holds it (kind "reference", one thread), else the OpenMP port.  Returns (cpu_baseline dict, ms per step, steps)."""
    from tools import synth
    db = np.concatenate([synth.sc_descriptors(min(10000, Kdb - s), first=s) for s in range(0, Kdb, 10000)])
    n_src = min(Kdb, 2000)
    src_rows = (np.arange(n_src, dtype=np.int64) * Kdb) // n_src
    qd, src, shift = synth.sc_queries(db[src_rows], sc_q)
    qd = qd[:Qc]
    use_ref = o.ref() is not None
    node = None
    try:
        if o.refsc() is not None:
            node = o.RefSCManager()                                 # the reference's own SCManager (include/Scancontext.cpp compiled unchanged)
            node.save_descriptors(db)                               # makeAndSaveScancontextAndKeys' bookkeeping on ready descriptors (:236-250)
    except Exception as e:
        print("[bench] reference SCManager unavailable (%r): falling back to the oracle port" % (e,), file=sys.stderr)
        node = None
    if node is None:
        keys = o.sc_keys_batch(db); qk = o.sc_keys_batch(qd)
    ts, tree_s = [], []
    for i in range(warm + steps):
        a = time.perf_counter()
        if node is not None:
            loop, sh = node.query_batch(qd); tb = 0.0               # every query through detectLoopClosureID (:253-344), unchanged; the tree is built in the batch's first call
        elif use_ref:
            loop, sh, dd, cand, (tb, tq) = o.ref_sc_query_batch(keys, db, qk, qd)
        else:
            loop, sh, dd, cand = o.sc_query_batch(keys, db, qk, qd); tb = 0.0
        dt = time.perf_counter() - a
        if i >= warm:
            ts.append(dt); tree_s.append(tb)
    v = Qc / float(np.median(ts))
    planted = src[:Qc] >= 0
    found = int(np.sum(loop[planted] == src_rows[src[:Qc][planted]]))
    if node is not None:
        kind, cores = "reference", 1
        sample = ("%d of the batch's %d queries per step, %d steps: every query through the reference's own SCManager::detectLoopClosureID (include/Scancontext.cpp compiled unchanged, "
                  "oracle/_ref/libliorf_ref_sc.so; Eigen = header stand-in with sequential reductions) against all %d keys — kd-tree (the reference's vendored nanoflann) built in the "
                  "step's first call, then 3 x distanceBtnScanContext per query; single-threaded as the reference's loop-closure thread is" % (Qc, sc_q, steps, Kdb))
    else:
        kind, cores = "port", all_cores
        sample = ("%d of the batch's %d queries per step, %d steps: %s over all %d keys built once per step (%.1f ms of a step) + distanceBtnScanContext of the 3 candidates, OpenMP over the queries"
                  % (Qc, sc_q, steps, "the reference's vendored nanoflann kd-tree" if use_ref else "brute-force top-3", Kdb, float(np.median(tree_s)) * 1e3))
    return dict(value=v, unit="queries/s", cores=cores, kind=kind, sample=sample, planted_loops_found=found, planted=int(planted.sum())), float(np.median(ts)) * 1e3, steps


def reference_rows(o, inst, kfs, threads, reps=5):
    """CPU figures beside the GPU line's `rows`: extractSurroundingKeyFrames of the 50 keyframes (extractNearby + transformPointCloud with OpenMP + the map VoxelGrid,
    src/mapOptmization.cpp:975-1044) on the reference's node, makeAndSaveScancontextAndKeys of the full scan and detectLoopClosureID on the reference's SCManager"""
    rows = {}
    if o.RefMapOpt.available() and os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libliorf_ref_mapopt_omp.so")):
        ts = []
        for r in range(reps + 1):
            R = o.RefMapOpt(openmp=True, numberOfCores=threads)     # a fresh node: the transformed-cloud cache (:1022-1032) starts empty, as for a new selection
            for k, (c, p) in enumerate(zip(kfs, inst["poses"])):
                R.add_keyframe(c, p.astype(np.float32), 100.0 + 0.1 * k)
            a = time.perf_counter()
            R.extract_surrounding_keyframes(100.0 + 0.1 * len(kfs))
            ts.append((time.perf_counter() - a) * 1e3)
            m = R.state()["m_ds"]
            R.close()
        rows["map_build_ms"] = dict(ms=stats_ms(ts[1:]), m_map=int(m), kind="reference", threads=threads,
                                    note="extractSurroundingKeyFrames on the reference's node: selection + 50 x transformPointCloud + VoxelGrid(0.5 m) of the concatenation")
    if o.refsc() is not None:
        S = o.RefSCManager()
        t_make, t_det = [], []
        for r in range(60):                                         # detectLoopClosureID needs more than 30 entries; the tree is rebuilt every 10th call
            a = time.perf_counter(); S.make_and_save(inst["scan"]); t_make.append((time.perf_counter() - a) * 1e3)
        for r in range(20):
            a = time.perf_counter(); S.detect(); t_det.append((time.perf_counter() - a) * 1e3)
        rows["sc_make_ms"] = dict(ms=stats_ms(t_make[5:]), n_points=int(len(inst["scan"])), kind="reference", threads=1)
        rows["sc_detect_ms"] = dict(ms=stats_ms(t_det), database=60, kind="reference", threads=1, note="one detectLoopClosureID call against 60 entries (tree rebuilt every 10th call)")
    return rows


def run_reference(args, W, K, world):
    """--impl reference: the CPU implementation of the path on the host cores — same metric / config as the GPU arm at this N, bounded sample."""
    sys.path.insert(0, ORACLE_DIR)
    import pyoracle as o
    all_cores = os.cpu_count() or 1
    o.set_num_threads(all_cores)                                  # torchrun exports OMP_NUM_THREADS=1: the baseline uses every host core regardless
    if world > 1 or args.gpus > 1:
        cb, ms_step, steps = sc_cpu_reference(o, args.sc_k, args.sc_q, 256, min(K, 10), min(W, 2), all_cores)
        v = cb["value"]
        line = dict(impl="reference", metric="sc_queries_per_s_100k", value=v, unit="queries/s", n_gpus=args.gpus, steps=steps, warmup=W, ms_per_step=ms_step,
                    higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32 ring keys (bf16 split filter + exact re-rank), f64 descriptors", data="synthetic", config=sc_config(args),
                    cpu_baseline=cb, e2e=dict(value=v, unit="queries/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return 0
    name = "kitti64_single"
    inst = make_single_inputs(name)
    _, _, ls, lm = SINGLE_CFGS[name]
    kfs = [o.voxel_grid(s, ls)[0] for s in inst["scans"]]
    mp, _, _ = o.voxel_grid(np.concatenate([o.transform_cloud(c, p.astype(np.float32)) for c, p in zip(kfs, inst["poses"])]), lm)
    flush = np.zeros(256 * 1024 * 1024 // 8)
    results = {}
    kind, sample, pose, n_ds = "port", "", None, 0
    for threads in sorted({all_cores, min(4, all_cores)}, reverse=True):
        cs = CpuStep(o, name, mp, threads)
        steps = K if threads == all_cores else min(K, 10)
        ts, split = [], {}
        for i in range(W + steps):
            flush += 1.0                                            # stream 256 MB through the caches between steps
            a = time.perf_counter()
            pose, tm, n_ds = cs.step(inst["scan"], inst["init"])
            c = time.perf_counter()
            if i >= W:
                ts.append((c - a) * 1e3)
                for k, v in tm.items():
                    split[k] = split.get(k, 0.0) + v / steps
        results[threads] = dict(ms=stats_ms(ts), steps=steps, split_ms=split)
        if threads == all_cores:
            kind, sample = cs.kind, cs.sample(name, steps, W)
        cs.close()
    o.set_num_threads(all_cores)
    best = results[all_cores]
    v = best["ms"]["mean"]
    line = dict(impl="reference", metric="scan2map_ms_per_frame_64beam", value=v, unit="ms/frame", n_gpus=args.gpus, steps=K, warmup=W, ms_per_step=v,
                higher_is_better=False, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", config=single_config(name, inst, n_ds, len(mp)),
                cpu_baseline=dict(value=v, unit="ms/frame", cores=all_cores, kind=kind, sample=sample,
                                  ms=best["ms"], split_ms=best["split_ms"], threads_4=results.get(4) if all_cores != 4 else None),
                e2e=dict(value=v, unit="ms/frame", h2d_bytes_per_step=0, d2h_bytes_per_step=0), final_pose=[float(x) for x in pose])
    try:                                                            # the other per-function figures BASELINE.md §3 lists, on the reference's own code where it is compiled
        line["rows"] = reference_rows(o, inst, kfs, all_cores)
    except Exception as e:
        line["rows"] = dict(error=repr(e))
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())

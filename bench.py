#!/usr/bin/env python
"""bench.py — headline benchmark of the liorf scan-to-map hot path on B200.

Workload (BASELINE.json configs[1], `kitti05_seq`): a synthetic 64-beam (HDL-64 shaped, ~119k returns/scan) driving
sequence; one STEP = one LiDAR frame through the whole path
    projectPointCloud (deskew) → downsampleCurrentScan → extractSurroundingKeyFrames (extractNearby selection +
    extractCloud + voxel-hash grid) → scan2MapOptimization (≤30 LM iterations, early exit as the reference) →
    saveFrame gate → keyframe store + ScanContext make, detectLoopClosureID every 10th frame.
`value`  : ms/frame with the raw scans already resident in HBM, timed with CUDA events on the library's stream.
`e2e`    : the same metric through the C ABI with HOST (pinned) raw scans: H2D of every scan and D2H of the pose inside
           the timed region.
Extra keys: `single_frame` (config 1: downsample + grid + 30 forced LM iterations against the local map — the "<1 ms"
target), `knn_queries_per_s`, `sc` (config 5: ScanContext queries/s over a K-entry database sharded over the ranks, exchanged
through NVLink peer-memory windows by the library's kernels), `batched`, `roofline`, `cpu_baseline`, `clocks`, `gpu_launches`.
Every frame announces its successor (liorf_frame_in.next): the next scan's H2D copy, deskew and downsample overlap this frame's solve.

`--impl reference` times the CPU path (oracle restatement + the reference's vendored nanoflann from oracle/_ref) on
the same frames, bounded sample, all host threads.
"""
import argparse
import math
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
ORACLE_DIR = os.path.join(ROOT, "oracle")     # imported ONLY by the cpu_baseline leg and --impl reference (never by the GPU arm)

T0 = 1000.0
DT = 0.1
KITTI = dict(lidarMinRange=1.0, lidarMaxRange=1000.0, N_SCAN=64, downsampleRate=2, point_filter_num=5)


# ----------------------------------------------------------------------------------------------------------------
# small SE(3) helpers (float64, host): initial guesses stand in for the IMU-preintegration output
# ----------------------------------------------------------------------------------------------------------------
def pose_to_T(p):
    r, pi_, y = p[0], p[1], p[2]
    A, B, C, D, E, F = np.cos(y), np.sin(y), np.cos(pi_), np.sin(pi_), np.cos(r), np.sin(r)
    T = np.eye(4)
    T[:3, :3] = [[A * C, A * D * F - B * E, B * F + A * D * E], [B * C, A * E + B * D * F, B * D * E - A * F], [-D, C * F, C * E]]
    T[:3, 3] = p[3:6]
    return T


def T_to_pose(T):
    return np.array([np.arctan2(T[2, 1], T[2, 2]), np.arcsin(-T[2, 0]), np.arctan2(T[1, 0], T[0, 0]), T[0, 3], T[1, 3], T[2, 3]])


class Sequence:
    """Seeded synthetic drive: poses, per-frame gyro tables and raw scans (generated lazily, cached)."""

    def __init__(self, n_frames, rank=0, filters=KITTI):
        from tools import synth
        self.synth = synth
        self.n = n_frames
        self.filters = filters
        self.poses = synth.street_trajectory(n_frames + 1, start=(0.0, 160.0 * rank), seed=synth.SEED0 + 1 + rank)
        self.rank = rank
        self.raw = {}
        self.imu = {}
        rng = np.random.default_rng(synth.SEED0 + 77 + rank)
        self.guess_noise = np.concatenate([rng.normal(scale=np.deg2rad(0.1), size=(n_frames, 3)), rng.normal(scale=0.02, size=(n_frames, 3))], axis=1)
        self.inc = [np.eye(4)] + [np.linalg.inv(pose_to_T(self.poses[i - 1])) @ pose_to_T(self.poses[i]) for i in range(1, n_frames)]
        self.inc_l = [tuple(tuple(float(v) for v in row) for row in m) for m in self.inc]
        self.noise_l = [tuple(float(v) for v in row) for row in self.guess_noise]

    def frame(self, i):
        if i not in self.raw:
            p = self.poses[i]
            omega = (self.poses[i + 1][:3] - p[:3]) / DT
            raw = self.synth.scan(self.synth.HDL64, p, omega=omega, vel=(0, 0, 0), seed=self.synth.SEED0 + 1000 * self.rank + i)
            t0 = T0 + DT * i
            it, rot, ptr = self.synth.imu_table(t0, t0 + float(raw["time"][-1]), omega, rate_hz=100.0, gyro_noise=1.56e-3, seed=i)
            self.raw[i] = raw
            self.imu[i] = (t0, it, rot, ptr)
            self.imu_cols = getattr(self, "imu_cols", {})
            self.imu_cols[i] = tuple(np.ascontiguousarray(rot[:, k]) for k in range(3))
        return self.raw[i], self.imu[i]

    def initial_guess(self, i, prev_est):
        """previous optimised pose ∘ true increment, perturbed by N(0, 0.1 deg / 2 cm).  Scalar float64 arithmetic (math module): this
        runs between two frames of the timed loop, on the critical path, so it must not cost tens of microseconds of numpy calls."""
        if i == 0 or prev_est is None:
            return self.poses[0].astype(np.float32)
        r, pi_, y = float(prev_est[0]), float(prev_est[1]), float(prev_est[2])
        A, Bs, Cc, D, E, F = math.cos(y), math.sin(y), math.cos(pi_), math.sin(pi_), math.cos(r), math.sin(r)
        R = ((A * Cc, A * D * F - Bs * E, Bs * F + A * D * E), (Bs * Cc, A * E + Bs * D * F, Bs * D * E - A * F), (-D, Cc * F, Cc * E))
        t = (float(prev_est[3]), float(prev_est[4]), float(prev_est[5]))
        M = self.inc_l[i]                                           # 4x4 increment as nested tuples
        # T = [R t] @ M : only the entries T_to_pose reads
        T00 = R[0][0] * M[0][0] + R[0][1] * M[1][0] + R[0][2] * M[2][0]
        T10 = R[1][0] * M[0][0] + R[1][1] * M[1][0] + R[1][2] * M[2][0]
        T20 = R[2][0] * M[0][0] + R[2][1] * M[1][0] + R[2][2] * M[2][0]
        T21 = R[2][0] * M[0][1] + R[2][1] * M[1][1] + R[2][2] * M[2][1]
        T22 = R[2][0] * M[0][2] + R[2][1] * M[1][2] + R[2][2] * M[2][2]
        tx = R[0][0] * M[0][3] + R[0][1] * M[1][3] + R[0][2] * M[2][3] + t[0]
        ty = R[1][0] * M[0][3] + R[1][1] * M[1][3] + R[1][2] * M[2][3] + t[1]
        tz = R[2][0] * M[0][3] + R[2][1] * M[1][3] + R[2][2] * M[2][3] + t[2]
        nz = self.noise_l[i]
        return (math.atan2(T21, T22) + nz[0], math.asin(-T20) + nz[1], math.atan2(T10, T00) + nz[2], tx + nz[3], ty + nz[4], tz + nz[5])


# ----------------------------------------------------------------------------------------------------------------
# one frame through the GPU library (the call sequence of cloudHandler + laserCloudInfoHandler)
# ----------------------------------------------------------------------------------------------------------------
class GpuPipeline:
    def __init__(self, seq, device):
        import liorf_b200
        self.ctx = liorf_b200.Context(device=device, **{k: seq.filters[k] for k in ("N_SCAN", "downsampleRate", "point_filter_num", "lidarMinRange", "lidarMaxRange")})
        self.ctx.reserve(131072, 4 << 20, 16 << 20, 4096)         # 180 GB of HBM: size once, never allocate inside a frame
        self.seq = seq
        self.prev = None
        self.stats = dict(frames=0, iters=0, knn_queries=0, alg_bytes_s2m=0, keyframes=0, n_ds=0, m_ds=0, loops=0)
        self.dev_raw = {}
        self.pin_raw = {}
        self.fin = {}

    def stage(self, frames):
        """raw scans → HBM (device arm) and pinned host memory (e2e arm), outside any timed region."""
        import torch
        for i in frames:
            raw, _ = self.seq.frame(i)
            if i not in self.dev_raw:
                t = torch.from_numpy(raw.view(np.uint8).reshape(-1).copy())
                self.pin_raw[i] = t.pin_memory()
                self.dev_raw[i] = self.pin_raw[i].to(f"cuda:{self.ctx.params.device}")
        torch.cuda.synchronize()
        for i in frames:
            self._frame_in(i, "dev"); self._frame_in(i, "e2e")

    def _src(self, i, mode):
        return (self.dev_raw[i].data_ptr(), True) if mode == "dev" else (self.pin_raw[i].data_ptr(), False)     # e2e: HOST (pinned) buffer, H2D inside the call

    def _frame_in(self, i, mode):
        """the frame's liorf_frame_in, built once (part of staging the inputs, like the scans themselves)"""
        fi = self.fin.get((i, mode))
        if fi is None:
            raw, (t0, it, rot, ptr) = self.seq.frame(i)
            p0, on0 = self._src(i, mode)
            fi = self.fin[(i, mode)] = self.ctx.frameIn(p0, len(raw), on0, t0, it, self.seq.imu_cols[i], ptr, True, loop_every=10, frame_index=i)
        return fi

    def step(self, i, mode, lookahead=True):
        """one frame = ONE call into the library (liorf_process_frame: the merged cloudHandler + laserCloudInfoHandler).
        lookahead: announce frame i+1 (liorf_frame_in.next) so that its H2D copy, deskew and downsample overlap this frame's solve."""
        guess = self.seq.initial_guess(i, self.prev)
        nxt = self._frame_in(i + 1, mode) if lookahead and (i + 1) in self.dev_raw else None
        fo = self.ctx.processFrameIn(self._frame_in(i, mode), guess, nxt)
        pose = np.array(fo.pose[:], np.float32)
        st = self.stats
        st["frames"] += 1; st["iters"] += fo.iters; st["knn_queries"] += fo.iters * max(fo.n_ds, 0)
        st["alg_bytes_s2m"] += 96 * fo.iters * max(fo.n_ds, 0); st["n_ds"] += max(fo.n_ds, 0); st["m_ds"] += max(fo.m_ds, 0)
        st["keyframes"] += fo.is_keyframe
        st["loops"] += int(fo.loop_checked and fo.loop_id >= 0)
        self.prev = pose
        return pose


# ----------------------------------------------------------------------------------------------------------------
# the CPU path (oracle + the reference's nanoflann) on the same frames — cpu_baseline and --impl reference
# ----------------------------------------------------------------------------------------------------------------
class CpuPipeline:
    def __init__(self, seq):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import pyoracle
        self.o = pyoracle
        self.seq = seq
        self.kf_clouds, self.kf_poses, self.kf_times = [], [], []
        self.prev = None
        self.state = np.zeros(37, np.float32)
        self.sc = pyoracle.SCManager()
        self.use_ref = pyoracle.ref() is not None
        self.split = dict(deskew=0.0, downsample=0.0, map_build=0.0, scan2map=0.0, sc=0.0)

    def seed_keyframes(self, clouds, poses, times):
        self.kf_clouds, self.kf_poses, self.kf_times = list(clouds), [np.asarray(p, np.float32) for p in poses], list(times)

    def extract_nearby(self, t_cur, radius=50.0, density=2.0):
        P = np.array(self.kf_poses, np.float32)[:, 3:6]
        d = ((P[-1] - P) ** 2).astype(np.float32).sum(1)
        near = np.lexsort((np.arange(len(P)), d))
        near = near[d[near] < radius * radius]
        ids = []
        last = P[-1]
        dist = lambda a, b: np.sqrt(((a - b) ** 2).astype(np.float32).sum(dtype=np.float32))
        if len(near):
            pts = np.concatenate([P[near], np.zeros((len(near), 1), np.float32)], 1)
            cent, _, _ = self.o.voxel_grid(pts, density)
            for c in cent:                                           # :1018 tests the voxel centroid, the id is its nearest real key pose
                if not dist(c[:3], last) > radius:
                    ids.append(int(np.argmin(((c[:3] - P) ** 2).sum(1))))
        for i in range(len(P) - 1, -1, -1):
            if t_cur - self.kf_times[i] < 10.0:
                if not dist(P[i], last) > radius:
                    ids.append(i)
            else:
                break
        return ids

    def step(self, i, guess=None, post=None):
        """guess: initial pose override (tests feed updateInitialGuess' output); post: callable applied to the solved pose
        (transformUpdate)"""
        o, seq = self.o, self.seq
        raw, (t0, it, rot, ptr) = seq.frame(i)
        a = time.perf_counter()
        cloud, _ = o.project_point_cloud(raw, seq.filters, t0, it, rot, ptr, True)
        b = time.perf_counter()
        ds, _, _ = o.voxel_grid(cloud, 0.4)
        c = time.perf_counter()
        if guess is None:
            guess = seq.initial_guess(i, self.prev)
        guess = np.asarray(guess, np.float32)
        pose = guess.copy()
        d = c
        if self.kf_clouds:
            ids = self.extract_nearby(t0)
            mraw = np.concatenate([o.transform_cloud(self.kf_clouds[k], self.kf_poses[k]) for k in ids], 0)
            mds, _, _ = o.voxel_grid(mraw, 0.5)
            d = time.perf_counter()
            self.last = dict(ds=ds, mds=mds, guess=guess.copy(), state=self.state.copy(), ids=list(ids))
            r = o.scan2map(ds, mds, guess, 30, False, self.state, use_ref_kdtree=self.use_ref)
            pose, self.state = r["tf"], r["state"]
            self.last["iters"] = r["iters"]
            if post is not None and r["iters"] > 0:
                pose = post(pose)
        e = time.perf_counter()
        last = self.kf_poses[-1] if self.kf_poses else None
        make = last is None
        if last is not None:
            Tb = np.linalg.inv(pose_to_T(last.astype(np.float64))) @ pose_to_T(pose.astype(np.float64))
            pb = T_to_pose(Tb)
            make = not (abs(pb[0]) < 0.2 and abs(pb[1]) < 0.2 and abs(pb[2]) < 0.2 and np.linalg.norm(pb[3:]) < 1.0)
        if make:
            self.kf_clouds.append(ds); self.kf_poses.append(pose.copy()); self.kf_times.append(t0)
            self.sc.make_and_save(cloud)
        if i % 10 == 9:
            self.sc.detect()
        f = time.perf_counter()
        s = self.split
        s["deskew"] += b - a; s["downsample"] += c - b; s["map_build"] += d - c; s["scan2map"] += e - d; s["sc"] += f - e
        self.prev = pose
        return pose


# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region by an in-process NVML thread (every ~2 ms: the timed
    region of the default run is only ~50-100 ms, too short for `nvidia-smi -lms`)."""

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = False
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(int(os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[gpu_index]) if os.environ.get("CUDA_VISIBLE_DEVICES", "").replace(",", "").isdigit() else gpu_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8), "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20), "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is None:
            return
        import threading
        self._stop = False
        self._thr = threading.Thread(target=self._loop, daemon=True)
        self._thr.start()

    def stop(self):
        if self.nv is None or self._thr is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvml unavailable"], samples=0)
        self._stop = True
        self._thr.join(timeout=2)
        sm = self.samples
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons), samples=len(sm))


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0), "fallback"


def dist_env():
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), ws


# ----------------------------------------------------------------------------------------------------------------
def bench_sc(ctx_device, rank, world, K, Q, reps, dist, peaks, q_large=0, sc_lanes=2):
    """config 5: K-entry database sharded by contiguous ranges over the ranks; Q replicated queries per step.
    One library call per batch and rank (liorf_b200/sc_sharded.py: PeerShardedSearch → liorf_sc_shard_query_dev): the exchange is done by
    the kernels through NVLink peer windows (csrc/sc_shard.cuh), the batch is replayed from a CUDA graph.  The ring-key stage runs on
    the tensor cores (csrc/sc_tensor.cuh); its GEMM kernel is timed live by the library's CUDA events in a few extra batches."""
    import torch
    import liorf_b200
    from liorf_b200.sc_sharded import GpuOps, ShardedScanContextSearch
    from tools import synth
    kloc = K // world
    off = rank * kloc
    ctx = liorf_b200.Context(device=ctx_device)
    ctx.reserve(1024, 1024, 0, kloc)
    CH = 10000
    for s in range(0, kloc, CH):
        ctx.scAddDescriptors(synth.sc_descriptors(min(CH, kloc - s), first=off + s))
    # queries: column-shifted noisy copies of 2000 database entries spread evenly over ALL rows (+ fresh ones); identical on every rank.
    # (Sources taken from the first rows only would put every true loop's candidates — and with them most of stage 2 — on rank 0.)
    n_src = min(K, 2000)
    src_rows = (np.arange(n_src, dtype=np.int64) * K) // n_src
    sample = np.concatenate([synth.sc_descriptors(1, first=int(i)) for i in src_rows])
    ops = GpuOps(ctx, off, torch)
    search = ShardedScanContextSearch(ops, rank, world, dist)      # world == 1: plain local search (no exchange at all)
    dev = ops.dev
    # exchange through NVLink peer windows (csrc/sc_shard.cuh), no NCCL inside a batch; one rank = the same code path with nothing to wait for
    from liorf_b200.sc_sharded import PeerShardedSearch
    peer = PeerShardedSearch(ctx, rank, world, [g * kloc for g in range(world + 1)], max(Q, q_large), torch)
    peer.connect_processes(dist)
    # LANES query batches in flight per GPU: a batch is a chain of ~25 small dependent kernels and 4 exchanges that leaves most SMs idle most
    # of the time; a second context on its own stream (and its own peer windows) searches the SAME database (liorf_sc_borrow_database, nothing
    # copied) and takes every other batch
    lanes = [(ctx, peer)]
    for _ in range(max(1, sc_lanes) - 1):
        c2 = liorf_b200.Context(device=ctx_device)
        c2.scBorrowDatabase(ctx)
        p2 = PeerShardedSearch(c2, rank, world, [g * kloc for g in range(world + 1)], max(Q, q_large), torch)
        p2.connect_processes(dist)
        lanes.append((c2, p2))

    def run(Qn, reps_n):
        qd, src, shift = synth.sc_queries(sample, Qn)
        with torch.cuda.stream(ops.stream):
            d_q = torch.from_numpy(qd).to(dev)
        torch.cuda.synchronize()                                   # every lane's stream reads d_q
        if world > 1:
            dist.barrier()                                         # ranks generate their (identical) queries at different speeds; a batch waits only seconds for a peer

        turn = [0]

        def one():
            k = turn[0] % len(lanes); turn[0] += 1
            return lanes[k][1].query(d_q)

        def sync_all():
            for c_, _ in lanes:
                c_.sync()
            torch.cuda.synchronize()
        lane_streams = [p_.stream for _, p_ in lanes]
        for _ in range(4 * len(lanes)):                            # warm-up (the third identical request of a lane captures its batch as a CUDA graph)
            loop, sh, dd, cand = one()
        sync_all()
        peer.wait_stats()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        turn[0] = 0
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ops.stream):
            e0.record()
        for st_ in lane_streams[1:]:
            st_.wait_event(e0)                                     # every lane starts inside the timed region
        n_batches = reps_n * len(lanes)
        for _ in range(n_batches):
            one()
        for st_ in lane_streams[1:]:
            ops.stream.wait_stream(st_)                            # ... and ends inside it
        with torch.cuda.stream(ops.stream):
            e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        waits, _ = peer.wait_stats()
        if world > 1:
            t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        # the tcgen05 GEMM's own duration: a few more batches of the same work with the library's CUDA-event sections on (plain launches)
        ctx.enableTiming(True)
        if world > 1:
            dist.barrier()
        n_timed = max(3, reps_n // 2)
        for _ in range(n_timed):
            loop, sh, dd, cand = lanes[0][1].query(d_q)            # lane 0 only: its context is the one with the sections on
        sync_all()
        tm = ctx.getTiming(); ctx.enableTiming(False)
        st = ctx.scTensorStats()
        lp = loop.cpu().numpy(); shn = sh.cpu().numpy()
        src = np.where(src >= 0, src_rows[np.maximum(src, 0)], -1)      # sample index → global database row
        ok = (lp == src) & (src >= 0)
        gemm_ms = tm["sc_gemm"][0] / max(tm["sc_gemm"][1], 1)
        kpad, qpad = (kloc + 127) // 128 * 128, (Qn + 255) // 256 * 256
        tflops = 2.0 * 64 * kpad * qpad / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
        res = dict(K=K, Q=Qn, shards=world, batches_in_flight=len(lanes), ms_per_batch=ms / n_batches, queries_per_s=Qn * n_batches / (ms * 1e-3),
                   planted_loops_found=int(ok.sum()), planted=int((src >= 0).sum()), shifts_correct=int((shn[ok] == shift[ok]).sum()),
                   ringkey_path="tcgen05 filter + exact re-rank" if tm["sc_gemm"][1] > 0 else "cuda-core brute force",
                   exchange=("NVLink peer windows (push + system-scope flags from the kernels, 4 phases per batch), no NCCL; batch replayed from a CUDA graph" if world > 1 else "none (one shard)"),
                   ringkey_stage_ms=tm["sc_search"][0] / n_timed, candidates_per_query=st["candidates"] / max(Qn, 1) * 32,
                   overflow_queries=st["overflow"],
                   peer_wait_us_per_batch={k: v / 1e3 / reps_n for k, v in waits.items()},      # lane 0      # rank 0: time its consumer kernels spent waiting for the peers' pushes
                   roofline=dict(kernel="k_sc_tensor", bound="tensor", achieved=tflops, peak=peaks.get("bf16_tflops"), unit="TFLOP/s",
                                 frac=(tflops / peaks["bf16_tflops"]) if tflops and peaks.get("bf16_tflops") else None, traffic=None,
                                 avg_launch_ms=gemm_ms,
                                 note="executed tensor-core flops: 2 x 64 (split-bf16 contraction) x Kpad x Qpad per launch; the distance itself is 3 x 20 flops per pair"))
        if world == 1:
            # stage 2 alone (distanceBtnScanContext for the 3 Q pairs of this batch), timed live: the HBM-bound kernel of the search.
            # Algorithmic bytes per pair: the candidate's 9 600-byte descriptor + a third of the query's (three pairs share it) + sector
            # keys and column norms of both (4 x 480 B).
            import ctypes as C
            with torch.cuda.stream(ops.stream):
                qq = ops.prepare_dev(d_q)
                pd_ = torch.empty((Qn, 3), dtype=torch.float64, device=dev); ps_ = torch.empty((Qn, 3), dtype=torch.int32, device=dev)
            vp = lambda t_: C.c_void_p(t_.data_ptr())
            call = lambda: ctx.lib.liorf_sc_distance_batch_dev(ctx.h, vp(d_q), vp(qq["sk"]), vp(qq["cn"]), vp(cand), Qn, 0, vp(pd_), vp(ps_))
            for _ in range(2):
                call()
            ctx.sync()
            s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(ops.stream):
                s0.record()
            for _ in range(5):
                call()
            with torch.cuda.stream(ops.stream):
                s1.record()
            ctx.sync()
            s2_ms = s0.elapsed_time(s1) / 5
            alg = 3 * Qn * (9600 + 9600 / 3 + 4 * 480)
            traffic2 = None
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
                traffic2 = tj.get("sc_distance_bulk", {}).get("dram_bytes_per_launch") if Qn == 32768 else None
            except Exception:
                pass
            res["stage2_roofline"] = dict(kernel="k_sc_distance_bulk", bound="hbm", achieved=alg / (s2_ms * 1e-3) / 1e9, peak=peaks["hbm_gbs"], unit="GB/s",
                                          frac=alg / (s2_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], traffic=traffic2, algorithmic_bytes_per_launch=alg, avg_launch_ms=s2_ms,
                                          pairs=3 * Qn, note="fp64 sums in the reference's sequential order: ~3 300 warp-instructions per pair, issue slots 54 % busy (ncu)")
        return res, qd
    res, qd = run(Q, reps)
    if q_large > Q:
        res["large_batch"], _ = run(q_large, max(2, reps // 2))
        res["large_batch"].pop("roofline", None)
    for c_, _ in lanes[1:]:
        c_.close()
    ctx.close()
    return res, (qd, sample)


def bench_batched(local_rank, rank, n_seq, P, W, K):
    """SURVEY §8d "batched figure": n_seq INDEPENDENT sequences in flight on one GPU (one context, host thread and pair of CUDA
    streams each; ctypes releases the GIL inside the library).  One sequence leaves the GPU idle while the host synchronises on
    the pose and between dependent launches; several fill those gaps.  Wall-clock aggregate, device-resident inputs."""
    import threading
    import torch
    n_frames = P + W + K + 1
    seqs = [Sequence(n_frames, 100 + 10 * rank + s) for s in range(n_seq)]
    for q in seqs:
        for i in range(n_frames):
            q.frame(i)
    pipes = [GpuPipeline(q, local_rank) for q in seqs]
    for p in pipes:
        p.stage(range(n_frames))
        for i in range(P + W):
            p.step(i, "dev")
    torch.cuda.synchronize()
    start = threading.Barrier(n_seq + 1)
    done = [0.0] * n_seq

    def run(k):
        start.wait()
        for i in range(P + W, P + W + K):
            pipes[k].step(i, "dev")
        pipes[k].ctx.sync()
        done[k] = time.perf_counter()
    th = [threading.Thread(target=run, args=(k,)) for k in range(n_seq)]
    for t in th:
        t.start()
    start.wait()
    t0 = time.perf_counter()
    for t in th:
        t.join()
    wall = max(done) - t0
    for p in pipes:
        p.ctx.close()
    return dict(sequences_in_flight=n_seq, frames=n_seq * K, ms_per_frame=wall * 1e3 / (n_seq * K), frames_per_s=n_seq * K / wall,
                timing="host wall clock around all threads (each ends with a stream synchronise)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=150)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--preroll", type=int, default=100, help="untimed frames that build the ~50-keyframe local map first")
    ap.add_argument("--sc-k", type=int, default=100000)
    ap.add_argument("--sc-q", type=int, default=4096)
    ap.add_argument("--sc-q-large", type=int, default=32768, help="second, larger query batch for the sharded search (0 = skip)")
    ap.add_argument("--batched", type=int, default=2, help="independent sequences in flight on one GPU for the batched figure (0 = skip)")
    ap.add_argument("--no-sc", action="store_true")
    ap.add_argument("--sc-lanes", type=int, default=0, help="ScanContext query batches in flight per GPU (contexts sharing one database); 0 = 2 on one GPU, 4 when sharded")
    ap.add_argument("--cpu-frames", type=int, default=12, help="bounded CPU-baseline sample (frames)")
    args = ap.parse_args()
    rank, local_rank, world = dist_env()
    W = max(args.warmup, 3)
    K = args.steps

    if args.impl == "reference":
        if rank != 0:
            return 0
        return run_reference(args, W, K)

    import torch
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    P = args.preroll
    B = min(K, 40)                                              # breakdown window after the timed one: every section timed (see below)
    n_frames = P + W + K + B + 1                                # + 1: the last frame announces (and pre-processes) its successor like every other
    seq = Sequence(n_frames, rank)
    for i in range(n_frames):                                   # synthesise everything up front (not timed)
        seq.frame(i)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    results = {}
    sampler = ClockSampler(local_rank)
    for mode in ("e2e", "dev"):
        pipe = GpuPipeline(seq, local_rank)
        pipe.stage(range(n_frames))
        for i in range(P):
            pipe.step(i, "dev")
        for i in range(P, P + W):
            pipe.step(i, mode)
        pipe.stats = {k: 0 for k in pipe.stats}
        # inside the timed window only the dominant kernel (the solver) is bracketed by CUDA events: every timed section costs two
        # cudaEventRecord calls of host time on the frame's critical path.  The per-section breakdown comes from the window after it.
        pipe.ctx.enableTiming(True, sections=["scan2map"])
        launches0 = pipe.ctx.launchCount()
        ext = torch.cuda.ExternalStream(pipe.ctx.stream(), device=dev)
        barrier()
        if mode == "dev" and rank == 0:                             # rank 0 samples: eight in-process NVML pollers contend on the driver and slow every rank's launches
            sampler.start()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        wall0 = time.perf_counter()
        with torch.cuda.stream(ext):
            e0.record()
        for i in range(P + W, P + W + K):
            pipe.step(i, mode)
        with torch.cuda.stream(ext):
            e1.record()
        barrier()
        wall = time.perf_counter() - wall0
        ms = e0.elapsed_time(e1)
        if mode == "dev":
            clocks = sampler.stop() if rank == 0 else None
        if world > 1:
            t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        results[mode] = dict(ms=ms, wall_ms=wall * 1e3, stats=dict(pipe.stats), timing=pipe.ctx.getTiming(), launches=pipe.ctx.launchCount() - launches0,
                             h2d=int(np.mean([seq.raw[i].nbytes for i in range(P + W, P + W + K)])) + 4 * 8 * 16 + 24)
        if mode == "dev":
            keep = pipe
        else:
            pipe.ctx.close()

    # ---- config 1: single-frame solve (downsample + grid build + 30 forced LM iterations) on the resident map ----
    pipe = keep
    ctx = pipe.ctx
    i_last = P + W + K - 1                                      # (the frame the single-frame case has always used: last of the timed window)
    raw, (t0, it, rot, ptr) = seq.frame(i_last)
    xyz = np.stack([raw["x"], raw["y"], raw["z"], raw["i"]], 1).astype(np.float32)      # filters off: all ~119k returns (BASELINE wording)
    ids = ctx.extractNearby(t0, 2.0)
    guess = (pipe.prev + np.array([np.deg2rad(0.5), np.deg2rad(0.3), np.deg2rad(1.5), 0.35, 0.1, 0.02], np.float32)).astype(np.float32)
    d_xyz = torch.from_numpy(xyz).to(dev)
    ext = torch.cuda.ExternalStream(ctx.stream(), device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    single = []
    ctx.enableTiming(True)
    kf0_pose = ctx.getKeyframe(0)[1]
    for rep in range(3 + 10):
        ctx.updateKeyframePose(0, kf0_pose)                       # invalidates the map cache → the map + grid are rebuilt every repetition
        ctx.setCurrentScanDev(d_xyz.data_ptr(), len(xyz))
        with torch.cuda.stream(ext):
            flush.zero_()                                         # L2 flush between timed iterations (256 MB > 126 MB L2)
        ctx.sync()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True); c = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ext):
            a.record()
        ctx.extractSurroundingKeyFrames(ids, want_count=False)
        with torch.cuda.stream(ext):
            b.record()
        ctx.downsampleCurrentScan(want_output=False)
        ctx.scan2MapOptimizationAsync(guess, 30, True)
        with torch.cuda.stream(ext):
            c.record()
        pose = ctx.getPose()
        if rep >= 3:
            single.append((a.elapsed_time(b), b.elapsed_time(c)))
    cnt = ctx.lastCounts()
    single = np.array(single)
    tsec = ctx.getTiming()
    s2m_ms = tsec["scan2map"][0] / max(tsec["scan2map"][1], 1)
    single_frame = dict(workload="kitti64_single: %d-pt scan, N_ds=%d, M=%d (%d keyframe clouds), 30 forced LM iterations" % (len(xyz), cnt["n_ds"], cnt["m_ds"], len(ids)),
                        solve_ms=float(np.median(single[:, 1])), solve_ms_p95=float(np.percentile(single[:, 1], 95)),
                        map_build_ms=(tsec["map_build"][0] + tsec["grid_build"][0]) / max(tsec["map_build"][1], 1),       # live CUDA events on the map stream (transform + VoxelGrid + grid)
                        solver_kernel_ms=s2m_ms,
                        knn_queries_per_s=30 * cnt["n_ds"] / (s2m_ms * 1e-3), target_ms=1.0)

    # ---- breakdown window: the next B frames of the same drive with EVERY section timed ----
    pipe.ctx.enableTiming(True)
    barrier()
    b0 = torch.cuda.Event(enable_timing=True); b1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ext):
        b0.record()
    for i in range(P + W + K, P + W + K + B):
        pipe.step(i, "dev")
    with torch.cuda.stream(ext):
        b1.record()
    barrier()
    results["breakdown"] = dict(ms=b0.elapsed_time(b1), frames=B, timing=pipe.ctx.getTiming())

    # ---- batched figure: several independent sequences in flight on this GPU ----
    batched = None
    if args.batched > 1:
        batched = bench_batched(local_rank, rank, args.batched, min(P, 60), W, min(K, 60))

    # ---- roofline of the dominant kernel of the sequence step ----
    peaks, peak_src = load_peaks()
    dv = results["dev"]
    bd = results["breakdown"]
    tb = bd["timing"]                                           # all sections, breakdown window
    tm = dict(tb); tm["scan2map"] = dv["timing"]["scan2map"]    # the solver: timed inside the headline window
    dom = max(("scan2map", "map_build", "downsample", "deskew", "grid_build"), key=lambda k: tb[k][0])
    if dom != "scan2map":                                       # only the solver is timed in the headline window; anything else falls back to the breakdown window
        tm = tb
    st = dv["stats"]
    alg = dict(scan2map=st["alg_bytes_s2m"],
               map_build=0, downsample=0, deskew=0, grid_build=0)
    n_raw = float(np.mean([len(seq.raw[i]) for i in range(P + W, P + W + K)]))
    alg["deskew"] = (24 * n_raw + 16 * n_raw / 10) * tm["deskew"][1]
    alg["downsample"] = (16 * n_raw / 10 + 16 * st["n_ds"] / max(st["frames"], 1)) * tm["downsample"][1]
    alg["map_build"] = (32 * 0 + 16 * st["m_ds"] / max(st["frames"], 1)) * tm["map_build"][1]       # + 16*M_raw in (added below when known)
    achieved = alg[dom] / (tm[dom][0] * 1e-3) / 1e9 if tm[dom][0] > 0 else 0.0
    traffic = None
    try:                                                        # dram bytes per launch from the committed `ncu --set full` capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj.get(dom, {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = dict(kernel=dom, bound="hbm", achieved=achieved, peak=peaks["hbm_gbs"], unit="GB/s", frac=achieved / peaks["hbm_gbs"], traffic=traffic,
                    algorithmic_bytes_per_launch=alg[dom] / max(tm[dom][1], 1),
                    peak_source=peak_src, share_of_step={k: tb[k][0] / bd["ms"] for k in tb}, avg_launch_ms=tm[dom][0] / max(tm[dom][1], 1),
                    timing="solver: CUDA events inside the timed window; share_of_step: the %d-frame window after it with every section timed (%.4f ms/frame)"
                           % (bd["frames"], bd["ms"] / bd["frames"]))

    # ---- ScanContext search (config 5) ----
    sc = None
    if not args.no_sc:
        sc, (qd, sample) = bench_sc(local_rank, rank, world, args.sc_k, args.sc_q, 10, dist, peaks, q_large=args.sc_q_large, sc_lanes=args.sc_lanes if args.sc_lanes > 0 else (2 if world == 1 else 4))

    # ---- CPU baseline (rank 0, N=1 only): the same frames on the host cores ----
    cpu = None
    if rank == 0 and world == 1 and args.cpu_frames > 0:
        cpu = run_cpu_sample(seq, keep, P, W, args.cpu_frames)
        if sc is not None:
            import pyoracle as o
            nq = 32
            kk = min(args.sc_k, 20000)
            from tools import synth
            db = synth.sc_descriptors(kk, first=0)
            keys = np.stack([o.sc_keys_from_desc(d)[0] for d in db])
            qk = np.stack([o.sc_keys_from_desc(d)[0] for d in qd[:nq]])
            t = time.perf_counter(); o.sc_query_batch(keys, db, qk, qd[:nq]); dt = time.perf_counter() - t
            cpu["sc_queries_per_s"] = nq / dt * (kk / args.sc_k)      # brute-force cost scales linearly in K
            cpu["sc_sample"] = f"{nq} queries x {kk}-entry database, scaled to K={args.sc_k}"
    keep.ctx.close()

    if rank == 0:
        tot_frames = K * world
        line = dict(metric="scan2map_ms_per_frame_64beam", value=dv["ms"] / tot_frames, unit="ms/frame", n_gpus=world, steps=K, warmup=W,
                    ms_per_step=dv["ms"] / K, higher_is_better=False, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                    config=dict(workload="kitti05_seq: synthetic 64-beam drive, %d-frame window after a %d-frame pre-roll, ~%d returns/scan, yaml filters (downsampleRate 2, point_filter_num 5), leaf 0.4/0.5, early-exit LM; one independent sequence per GPU"
                                % (K, P, int(n_raw)), l2="inputs streamed: every frame reads a fresh 2.9 MB scan; per-frame working set is not reused across frames",
                                avg_n_ds=st["n_ds"] / max(st["frames"], 1), avg_m_ds=st["m_ds"] / max(st["frames"], 1), avg_lm_iters=st["iters"] / max(st["frames"], 1),
                                keyframes_added=st["keyframes"],
                                pipeline="liorf_frame_in.next: frame i+1's H2D copy, deskew and downsample run on a second stream while frame i is solved (the reference's imageProjection / mapOptimization node pair); results bit-identical to the unpipelined call"),
                    wall_ms_per_step=dv["wall_ms"] / K,
                    e2e=dict(value=results["e2e"]["ms"] / tot_frames, unit="ms/frame", h2d_bytes_per_step=results["e2e"]["h2d"], d2h_bytes_per_step=64,
                             wall_ms_per_step=results["e2e"]["wall_ms"] / K),
                    gpu_launches=dv["launches"], knn_queries_per_s=st["knn_queries"] / (tm["scan2map"][0] * 1e-3) if tm["scan2map"][0] > 0 else None,
                    single_frame=single_frame, batched=batched, sc=sc, roofline=roofline, cpu_baseline=cpu, clocks=clocks,
                    kernel_ms_per_frame={k: tb[k][0] / bd["frames"] for k in tb})
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_cpu_sample(seq, gpu_pipe, P, W, n_frames):
    """CPU path on frames [P+W, P+W+n) starting from the keyframe state the GPU run had at frame P+W (same inputs)."""
    if ORACLE_DIR not in sys.path:
        sys.path.insert(0, ORACLE_DIR)
    import pyoracle as o
    cpu = CpuPipeline(seq)
    # rebuild the state as of frame P+W by replaying the GPU context's keyframes that existed then
    ctx = gpu_pipe.ctx
    clouds, poses, times = [], [], []
    t_cut = T0 + DT * (P + W)
    for k in range(ctx.numKeyframes()):
        cl, ps, tt = ctx.getKeyframe(k)
        if tt < t_cut - 1e-9:
            clouds.append(cl); poses.append(ps); times.append(tt)
    cpu.seed_keyframes(clouds, poses, times)
    for cl in clouds[-40:]:
        cpu.sc.save_descriptor(np.zeros(1200))
    cpu.prev = poses[-1] if poses else None
    cpu.step(P + W)                                              # warm-up frame
    cpu.split = {k: 0.0 for k in cpu.split}
    t = time.perf_counter()
    for i in range(P + W + 1, P + W + 1 + n_frames):
        cpu.step(i)
    dt = time.perf_counter() - t
    return dict(value=dt / n_frames * 1e3, unit="ms/frame", cores=o.num_threads(), kind="port",
                sample=f"{n_frames} consecutive frames of the same sequence from the same keyframe state; oracle restatement, kd-tree = the reference's vendored nanoflann"
                       + ("" if cpu.use_ref else " (oracle/_ref missing: brute-force kNN)"),
                split_ms_per_frame={k: v / n_frames * 1e3 for k, v in cpu.split.items()})


def run_reference(args, W, K):
    """--impl reference: the CPU implementation of the path on the host cores, same metric/config, bounded sample."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as o
    P = min(args.preroll, 60)
    steps = min(K, 20)
    seq = Sequence(P + W + steps, 0)
    for i in range(P + W + steps):                               # synthesise the scans up front: generation is not part of the path
        seq.frame(i)
    cpu = CpuPipeline(seq)
    for i in range(P + W):
        cpu.step(i)
    cpu.split = {k: 0.0 for k in cpu.split}
    t = time.perf_counter()
    for i in range(P + W, P + W + steps):
        cpu.step(i)
    dt = time.perf_counter() - t
    v = dt / steps * 1e3
    line = dict(impl="reference", metric="scan2map_ms_per_frame_64beam", value=v, unit="ms/frame", n_gpus=args.gpus, steps=steps, warmup=W, ms_per_step=v,
                higher_is_better=False, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload="kitti05_seq: synthetic 64-beam drive, CPU path (oracle restatement + the reference's vendored nanoflann kd-tree), "
                                     "%d timed frames after a %d-frame pre-roll" % (steps, P + W)),
                cpu_baseline=dict(value=v, unit="ms/frame", cores=o.num_threads(), kind="port",
                                  sample=f"{steps} consecutive frames", split_ms_per_frame={k: x / steps * 1e3 for k, x in cpu.split.items()}),
                e2e=dict(value=v, unit="ms/frame", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())

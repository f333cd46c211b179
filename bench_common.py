"""bench_common.py — shared pieces of bench.py: the synthetic drive (Sequence), one frame through the GPU library (GpuPipeline) and through
the CPU path (CpuPipeline: oracle restatement + the reference's vendored nanoflann, test / baseline infrastructure), the NVML clock
sampler and the measured peaks.  Nothing here is timed on its own."""
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
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

    def __init__(self, n_frames, rank=0, filters=KITTI, loop=False):
        from tools import synth
        self.synth = synth
        self.n = n_frames
        self.filters = filters
        self.poses = synth.street_trajectory(n_frames + 1, start=(0.0, 160.0 * rank), seed=synth.SEED0 + 1 + rank, loop=loop)
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



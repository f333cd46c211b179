"""Seeded synthetic inputs: ray-cast scans, trajectories, gyro tables, ScanContext databases.
Test / benchmark infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PRAW = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("i", "<f4"), ("ring", "<u2"), ("pad", "<u2"), ("time", "<f4")])
SEED0 = 20240517
HDL64, OS1_128, LIVOX = 0, 1, 2
_lib = None


def build():
    so = os.path.join(HERE, "libliorf_synth.so")
    src = os.path.join(HERE, "synth.cpp")
    if not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.run(["make", "-C", HERE, "-s"], check=True)
    return so


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def scan(sensor, pose6, omega=(0, 0, 0), vel=(0, 0, 0), seed=SEED0, noise=0.02):
    """pose6 = (roll,pitch,yaw,x,y,z) world pose of the sensor at scan start.  Returns a PRAW structured array."""
    cap = lib().synth_scan_capacity(C.c_int(sensor))
    out = np.zeros(cap, PRAW)
    p = np.ascontiguousarray(pose6, np.float64); w = np.ascontiguousarray(omega, np.float64); v = np.ascontiguousarray(vel, np.float64)
    n = lib().synth_scan(C.c_int(sensor), p.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p),
                         C.c_uint64(seed), C.c_double(noise), out.ctypes.data_as(C.c_void_p), C.c_int(cap))
    return out[:n].copy()


def raw_to_xyzi(raw):
    return np.stack([raw["x"], raw["y"], raw["z"], raw["i"]], axis=1).astype(np.float32)


def sc_descriptors(count, first=0, seed=SEED0 + 4):
    d = np.zeros((count, 1200), np.float64)
    lib().synth_sc_descriptors(C.c_uint64(seed), C.c_int(first), C.c_int(count), d.ctypes.data_as(C.c_void_p))
    return d


def sc_queries(db, Q, seed=SEED0 + 4):
    db = np.ascontiguousarray(db, np.float64).reshape(-1, 1200)
    q = np.zeros((Q, 1200), np.float64); src = np.zeros(Q, np.int32); shift = np.zeros(Q, np.int32)
    lib().synth_sc_queries(C.c_uint64(seed), db.ctypes.data_as(C.c_void_p), C.c_int(len(db)), C.c_int(Q), q.ctypes.data_as(C.c_void_p),
                           src.ctypes.data_as(C.c_void_p), shift.ctypes.data_as(C.c_void_p))
    return q, src, shift


def imu_table(time_scan_cur, time_scan_end, omega, rate_hz=100.0, gyro_noise=0.0, seed=SEED0):
    """The table imuDeskewInfo (src/imageProjection.cpp:350-409) builds from a gyro stream of constant (or noisy) body rate:
    row 0 = zeros at the first stamp >= Cur - 0.01, then rectangle-rule integration, stop after End + 0.01.
    Returns (imu_time[rows], imu_rot[rows,3], imu_pointer_cur)."""
    rng = np.random.default_rng(seed)
    dt = 1.0 / rate_hz
    k0 = int(np.ceil((time_scan_cur - 0.01) / dt - 1e-9))
    times, rots = [], []
    k = k0
    while True:
        t = k * dt
        if t > time_scan_end + 0.01:
            break
        w = np.asarray(omega, np.float64) + (rng.normal(scale=gyro_noise, size=3) if gyro_noise > 0 else 0.0)
        if not times:
            times.append(t); rots.append(np.zeros(3))
        else:
            td = t - times[-1]
            rots.append(rots[-1] + w * td); times.append(t)
        k += 1
    return np.array(times, np.float64), np.array(rots, np.float64).reshape(-1, 3), len(times) - 1


def street_trajectory(n_frames, speed=8.0, dt=0.1, start=(0.0, 0.0), seed=SEED0 + 1, loop=False):
    """Sensor poses (roll,pitch,yaw,x,y,z) along street centre lines of the 40 m lattice: straight runs with 90-degree
    turns at intersections.  With loop=True the path is a closed rectangle that revisits its first leg (loop closure)."""
    rng = np.random.default_rng(seed)
    poses = np.zeros((n_frames, 6), np.float64)
    x, y = start
    heading = 0.0
    leg_left = 200.0 if loop else 1e18
    turn_left = 0.0
    for i in range(n_frames):
        poses[i] = (0.002 * np.sin(0.13 * i), 0.003 * np.sin(0.07 * i + 1.0), heading, x, y, 0.02 * np.sin(0.05 * i))
        v = speed + 2.0 * np.sin(0.021 * i + rng.uniform(0, 1e-9))
        step = v * dt
        if turn_left > 0:                       # turning through the intersection on a small arc
            dpsi = min(turn_left, 0.4 * dt * 2.0)
            heading += dpsi; turn_left -= dpsi
        x += step * np.cos(heading); y += step * np.sin(heading)
        leg_left -= step
        if leg_left <= 0 and turn_left <= 0:
            turn_left = np.pi / 2; leg_left = 200.0
    return poses

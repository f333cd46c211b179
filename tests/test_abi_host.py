"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/liorf_b200.h declares, refuses
to run without a device (no CPU fallback), and its host-side keyframe logic (extractNearby / saveFrame) behaves as the
reference's."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import liorf_b200
    liorf_b200.build_library()
    return liorf_b200.load_library()


def test_every_declared_symbol_is_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "liorf_b200.h")).read()
    names = list(dict.fromkeys(re.findall(r"\b(liorf_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 45
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_struct_layouts_match_header():
    from liorf_b200.api import Params, LMTrace, PRAW, P4
    assert C.sizeof(Params) == 12 * 4
    assert C.sizeof(LMTrace) == 64 * 6 * 4 + 64 * 4 + 4 * 4
    assert PRAW.itemsize == 24 and P4.itemsize == 16


def test_sm100a_code_only():
    """the library carries sm_100a SASS and nothing else (no multi-arch fatbin, no PTX JIT path)."""
    import subprocess
    import liorf_b200
    out = subprocess.run(["cuobjdump", "-lelf", liorf_b200.library_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, out


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from liorf_b200.api import Params
    p = Params.default()
    h = C.c_void_p()
    assert lib.liorf_create(C.byref(p), C.byref(h)) < 0 and not h.value      # fails loudly, nothing to fall back to


def test_package_does_not_import_oracle():
    import subprocess
    import sys
    code = "import sys; import liorf_b200; assert not any('pyoracle' in m or m.startswith('oracle') for m in sys.modules), 'oracle leaked'"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
    for root, _, files in os.walk(os.path.join(ROOT, "liorf_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                txt = open(os.path.join(root, f), errors="ignore").read()
                assert "pyoracle" not in txt and "liorf_oracle" not in txt, f


def _extract_nearby(lib, poses, times, t_cur, radius=50.0, density=2.0):
    poses = np.ascontiguousarray(poses, np.float32); times = np.ascontiguousarray(times, np.float64)
    ids = np.zeros(4096, np.int32); n = C.c_int(0)
    rc = lib.liorf_host_extract_nearby(poses.ctypes.data_as(C.c_void_p), times.ctypes.data_as(C.c_void_p), C.c_int(len(poses)), C.c_double(t_cur),
                                       C.c_float(radius), C.c_float(density), ids.ctypes.data_as(C.c_void_p), C.c_int(4096), C.byref(n))
    assert rc == 0
    return ids[:n.value].copy()


def test_extract_nearby_selection(lib):
    # 120 keyframes, 1 m apart along x, 0.2 s apart
    n = 120
    poses = np.zeros((n, 6), np.float32); poses[:, 3] = np.arange(n); times = 100.0 + 0.2 * np.arange(n)
    ids = _extract_nearby(lib, poses, times, times[-1] + 0.1)
    # radius part: poses within 50 m of the newest, thinned to 2 m voxels → about 25, each a real keyframe id
    recent = [i for i in range(n - 1, -1, -1) if times[-1] + 0.1 - times[i] < 10.0]
    n_recent = len(recent)
    head, tail = ids[:-n_recent], ids[-n_recent:]
    assert list(tail) == recent                                              # "last 10 s" appended newest first (:1000-1007)
    assert 20 <= len(head) <= 30 and np.all(np.diff(np.sort(head)) >= 1)
    assert np.all(poses[head, 3] > poses[-1, 3] - 50.0)
    assert set(head) & set(tail)                                             # duplicates are possible and kept (trap 10)
    # single keyframe
    assert list(_extract_nearby(lib, poses[:1], times[:1], times[0])) == [0, 0]
    # nothing recent: only the radius part remains
    ids2 = _extract_nearby(lib, poses, times, times[-1] + 100.0)
    assert len(ids2) == len(head) and np.array_equal(ids2, head)


def test_save_frame_gate(lib):
    f = lambda last, cur: lib.liorf_host_save_frame(None if last is None else np.asarray(last, np.float32).ctypes.data_as(C.c_void_p),
                                                    np.asarray(cur, np.float32).ctypes.data_as(C.c_void_p), C.c_float(1.0), C.c_float(0.2))
    assert f(None, [0, 0, 0, 0, 0, 0]) == 1                                   # first frame is always a keyframe
    assert f([0, 0, 0, 0, 0, 0], [0, 0, 0, 0.5, 0.5, 0]) == 0                  # 0.71 m < 1 m
    assert f([0, 0, 0, 0, 0, 0], [0, 0, 0, 0.8, 0.7, 0]) == 1                  # 1.06 m
    assert f([0, 0, 0, 0, 0, 0], [0, 0, 0.25, 0, 0, 0]) == 1                   # yaw 0.25 rad > 0.2
    assert f([0, 0, 1.0, 5, 5, 0], [0, 0, 1.1, 5.3, 5.3, 0]) == 0              # relative motion evaluated in the last keyframe's frame
    p = np.array([5.0, -3.0, 0.1, 1, 2, 12.0], np.float32)
    lib.liorf_transform_update_clamp(p.ctypes.data_as(C.c_void_p), C.c_float(1.0), C.c_float(10.0))
    assert list(p) == [1.0, -1.0, np.float32(0.1), 1.0, 2.0, 10.0]


# ---------------------------------------------------------------------------------------------- §8f-2 scalar pre / post steps
def _guess_lib(lib, st, no_kf, ci, tf, heading=True, imu_type=1):
    from liorf_b200 import CloudInfoGuess
    c = CloudInfoGuess(int(ci[0]), int(ci[1]), *[float(v) for v in ci[2:]])
    t = np.array(tf, np.float32)
    assert lib.liorf_host_update_initial_guess(C.byref(st), int(no_kf), C.byref(c), int(heading), int(imu_type), t.ctypes.data_as(C.c_void_p)) == 0
    return t


def test_update_initial_guess_matches_oracle_sequence(lib):
    """updateInitialGuess (src/mapOptmization.cpp:899-958) over a synthetic drive: first frame (no keyframes), first frame with
    odometry (only latches lastImuPreTransformation and falls through to the IMU-rotation branch), then pre-integration
    increments, with and without odometry / IMU — library host code vs the oracle's separate restatement, float for float."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as o
    from liorf_b200 import GuessState
    rng = np.random.default_rng(3)
    for imu_type, heading in ((1, True), (0, True), (1, False)):
        st = GuessState(); ost = np.zeros(31, np.float32); ost[6:18] = [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0]; ost[18:30] = ost[6:18]
        tf = np.zeros(6, np.float32)
        odo = np.array([0, 0, 0, 0.01, -0.02, 0.3], np.float64)      # x y z roll pitch yaw of the pre-integration odometry
        for k in range(40):
            odo += np.array([0.8, 0.05, 0.01, 0.001, -0.002, 0.02]) + rng.normal(scale=1e-3, size=6)
            imu_rpy = odo[3:] + rng.normal(scale=2e-3, size=3)
            ci = np.array([k % 7 != 6, k % 5 != 4, *imu_rpy, *odo], np.float32)
            no_kf = k == 0
            tf = _guess_lib(lib, st, no_kf, ci, tf, heading, imu_type)
            ost[:6] = ost[:6] if k else 0
            ost = o.update_initial_guess(ost, no_kf, ci, heading, imu_type)
            assert np.array_equal(tf, ost[:6]), (imu_type, heading, k, tf, ost[:6])
            assert np.array_equal(np.array(st.lastImuTransformation[:], np.float32), ost[6:18])
            assert np.array_equal(np.array(st.lastImuPreTransformation[:], np.float32), ost[18:30]) and st.lastImuPreTransAvailable == int(ost[30])
        assert tf[3] > 15.0 and tf[2] > 0.4                      # the chained increments carried the pose along the drive


def test_update_initial_guess_increment_is_body_frame(lib):
    from liorf_b200 import GuessState
    st = GuessState()
    tf = _guess_lib(lib, st, True, [1, 1, 0, 0, 0.5, 10, 20, 0, 0, 0, 0.5], np.zeros(6))     # first frame: attitude from the IMU
    assert np.allclose(tf, [0, 0, 0.5, 0, 0, 0])
    tf = _guess_lib(lib, st, False, [1, 1, 0, 0, 0.5, 10, 20, 0, 0, 0, 0.5], tf)            # latches the odometry pose, no motion
    assert np.allclose(tf, [0, 0, 0.5, 0, 0, 0], atol=1e-6)
    tf[2] = 1.0                                                                              # the optimiser turned the pose to yaw = 1.0
    c, s = np.cos(0.5), np.sin(0.5)
    tf = _guess_lib(lib, st, False, [1, 1, 0, 0, 0.5, 10 + c, 20 + s, 0, 0, 0, 0.5], tf)    # odometry: 1 m forward in ITS heading
    assert np.allclose(tf[3:5], [np.cos(1.0), np.sin(1.0)], atol=1e-5) and abs(tf[2] - 1.0) < 1e-6   # applied in the pose's own frame


def test_transform_update_slerp_and_clamps(lib):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as o
    rng = np.random.default_rng(4)
    for _ in range(200):
        tf = rng.normal(scale=[0.2, 0.2, 1.0, 30, 30, 3], size=6).astype(np.float32)
        ir, ip, w = float(rng.normal(scale=0.2)), float(rng.normal(scale=0.2)), float(rng.choice([0.0, 0.01, 0.3, 1.0]))
        avail, itype = int(rng.integers(0, 2)), int(rng.integers(0, 2))
        g = tf.copy()
        lib.liorf_host_transform_update(g.ctypes.data_as(C.c_void_p), avail, itype, C.c_float(ir), C.c_float(ip), C.c_float(w), C.c_float(0.3), C.c_float(2.0))
        r = o.transform_update(tf, avail, itype, ir, ip, w, 0.3, 2.0)
        assert np.array_equal(g, r)
        if avail and itype:                                        # single-axis slerp == linear blend of the angle
            assert abs(g[0] - np.clip((1 - w) * tf[0] + w * np.float32(ir), -0.3, 0.3)) < 2e-6
            assert abs(g[1] - np.clip((1 - w) * tf[1] + w * np.float32(ip), -0.3, 0.3)) < 2e-6
        else:
            assert g[0] == np.clip(tf[0], -0.3, 0.3) and g[1] == np.clip(tf[1], -0.3, 0.3)
        assert g[5] == np.clip(tf[5], -2.0, 2.0) and np.array_equal(g[2:5], tf[2:5])


def test_null_context_is_an_error_code_never_a_crash():
    """Error behaviour of the boundary: every entry that takes the context answers a NULL context with LIORF_ERR_ARG (-1 for the launch counter, NULL
    for the stream) — no CUDA call, no dereference.  Run in a child process so that a crash would fail the test instead of the suite."""
    import subprocess
    import sys
    code = r'''
import ctypes as C, re, sys
sys.path.insert(0, %r)
import liorf_b200
lib = liorf_b200.load_library()
hdr = re.sub(r"/\*.*?\*/", "", open(%r).read(), flags=re.S)
n_checked = 0
for ret, name, args in re.findall(r"\b(int|void|long long|const char\*|void\*)\s+(liorf_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
    args = " ".join(args.split())
    if not args.startswith("liorf_ctx*"):
        continue
    n = len(re.split(r",(?![^()]*\))", args))
    f = getattr(lib, name)
    f.restype = None if ret == "void" else (C.c_void_p if ret == "void*" else (C.c_longlong if ret == "long long" else C.c_int))
    r = f(*([C.c_void_p(0)] * n))
    if ret == "int":
        assert r == -2, (name, r)                     # LIORF_ERR_ARG
    elif ret == "long long":
        assert r == -1, (name, r)
    elif ret == "void*":
        assert not r, (name, r)
    n_checked += 1
assert n_checked >= 70, n_checked
# the context-free host entries: NULL mandatory pointers
for name, n in (("liorf_host_extract_nearby", 9), ("liorf_host_update_initial_guess", 6), ("liorf_host_transform_update", 8), ("liorf_host_save_frame", 4),
                ("liorf_host_imu_deskew_info", 14), ("liorf_create", 2)):
    f = getattr(lib, name); f.restype = C.c_int
    assert f(*([C.c_void_p(0)] * n)) == -2, name
for name, n in (("liorf_transform_update_clamp", 3), ("liorf_default_params", 1), ("liorf_destroy", 1)):
    f = getattr(lib, name); f.restype = None
    f(*([C.c_void_p(0)] * n))                        # void: must simply return
print("checked", n_checked)
''' % (ROOT, os.path.join(ROOT, "include", "liorf_b200.h"))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "checked" in p.stdout, (p.returncode, p.stdout[-500:], p.stderr[-1500:])

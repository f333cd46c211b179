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

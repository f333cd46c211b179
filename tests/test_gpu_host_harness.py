"""The C++ host mirror (liorf_b200/host/liorf_host.hpp: same member-function names as the reference) drives the same C ABI
as the Python binding: identical bytes in → identical pose out."""
import os
import struct
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_harness_matches_python_binding(ctx, kitti_case, tmp_path):
    exe = os.path.join(ROOT, "liorf_b200", "host", "liorf_harness")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.dirname(exe)], check=True)
    kfs = kitti_case["keyframes"]; scan = kitti_case["scan"]; init = kitti_case["init"]
    t_cur = 0.1 * len(kfs)
    p = tmp_path / "case.bin"
    with open(p, "wb") as f:
        f.write(struct.pack("<i", len(kfs)))
        for k, (c, pose) in enumerate(kfs):
            f.write(struct.pack("<i", len(c))); f.write(np.asarray(pose, np.float32).tobytes()); f.write(struct.pack("<d", 0.1 * k)); f.write(np.ascontiguousarray(c, np.float32).tobytes())
        f.write(struct.pack("<i", len(scan))); f.write(np.ascontiguousarray(scan, np.float32).tobytes())
        f.write(np.asarray(init, np.float32).tobytes()); f.write(struct.pack("<d", t_cur))
    out = subprocess.run([exe, str(p)], capture_output=True, text=True, check=True).stdout.split()
    cpp_pose = np.array([float(x) for x in out[1:7]], np.float32)
    for k, (c, pose) in enumerate(kfs):
        ctx.addKeyframeCloud(c, pose, 0.1 * k)
    ctx.setCurrentScan(scan)
    ctx.extractSurroundingKeyFrames(ctx.extractNearby(t_cur, 2.0), want_count=False)
    ctx.downsampleCurrentScan(want_output=False)
    pose, tr = ctx.scan2MapOptimization(init, 30, False)
    assert np.array_equal(cpp_pose, pose)                      # same library, same bytes → bit-identical
    assert int(out[8]) == tr.iters and int(out[10]) == tr.converged == 1
    assert np.linalg.norm(pose[3:] - kitti_case["truth"][3:]) < 0.05

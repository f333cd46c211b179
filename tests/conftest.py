import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # one hardware channel per stream; must precede the first CUDA call

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def synth():
    from tools import synth as s
    s.build()
    return s


@pytest.fixture()
def ctx():
    import liorf_b200
    c = liorf_b200.Context()
    yield c
    c.close()


@pytest.fixture(scope="session")
def kitti_case(synth, oracle):
    """Small version of config 1: HDL-64 scans along a street, 12 keyframes at 1 m spacing, 0.4 / 0.5 m leaves."""
    poses = [np.array([0, 0, 0, 1.0 * k, 0, 0], np.float64) for k in range(12)]
    kfs = []
    for k, p in enumerate(poses):
        raw = synth.scan(synth.HDL64, p, seed=synth.SEED0 + k)
        ds, _, _ = oracle.voxel_grid(synth.raw_to_xyzi(raw), 0.4)
        kfs.append((ds, p.astype(np.float32)))
    q_pose = np.array([0, 0, 0, 11.0, 0, 0], np.float64)
    raw = synth.scan(synth.HDL64, q_pose, seed=synth.SEED0 + 100)
    scan = synth.raw_to_xyzi(raw)
    init = (q_pose + np.array([np.deg2rad(0.5), np.deg2rad(0.3), np.deg2rad(1.5), 0.35, 0.1, 0.02])).astype(np.float32)
    return dict(keyframes=kfs, scan=scan, raw=raw, init=init, truth=q_pose.astype(np.float32))

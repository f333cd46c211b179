"""Loop-closure registration: the ICP of performSCLoopClosure (src/mapOptmization.cpp:624-730) on the GPU vs the
numpy/scipy-free CPU oracle in oracle/pyicp.py (brute-force neighbours, numpy SVD — a different implementation on purpose)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def _rot_err(Ta, Tb):
    R = Ta[:3, :3].astype(np.float64) @ Tb[:3, :3].astype(np.float64).T
    return float(np.arccos(np.clip((np.trace(R) - 1) / 2, -1, 1)))


@pytest.fixture(scope="module")
def revisit(synth, oracle):
    """12 keyframes along a street, then a 13th that revisits keyframe 4's place with an odometry drift of (0.6 m, -0.4 m, 2 deg)."""
    kfs = []
    for k in range(12):
        p = np.array([0, 0, 0, 1.0 * k, 0, 0], np.float64)
        ds, _, _ = oracle.voxel_grid(synth.raw_to_xyzi(synth.scan(synth.HDL64, p, seed=700 + k)), 0.4)
        kfs.append((ds, p.astype(np.float32)))
    true = np.array([0, 0, 0.0, 4.0, 0, 0], np.float64)
    ds, _, _ = oracle.voxel_grid(synth.raw_to_xyzi(synth.scan(synth.HDL64, true, seed=799)), 0.4)
    drift = np.array([0, 0, np.deg2rad(2.0), 4.6, -0.4, 0.0], np.float32)
    kfs.append((ds, drift))
    return kfs


def test_icp_matches_oracle(oracle, synth, revisit):
    import liorf_b200
    import pyicp
    c = liorf_b200.Context()
    for cl, p in revisit:
        c.addKeyframeCloud(cl, p)
    cur, pre = len(revisit) - 1, 4
    # loop_index = -1: every cloud by its own pose (performRSLoopClosure's form) so that the drifted pose matters
    r = c.loopClosureICP(cur, pre, history_search_num=2, loop_index=-1, icp_leaf=0.5, max_corr_dist=20.0, max_iters=100)
    assert r.ran == 1
    src, tgt = c.icpClouds(r.n_source, r.n_target)
    clouds, poses = [k[0] for k in revisit], [k[1] for k in revisit]
    o_src = pyicp.loop_find_near_keyframes(clouds, poses, cur, 0, -1, 0.5, oracle)
    o_tgt = pyicp.loop_find_near_keyframes(clouds, poses, pre, 2, -1, 0.5, oracle)
    assert np.array_equal(src, o_src) and np.array_equal(tgt, o_tgt)               # the two clouds: bit-exact
    o = pyicp.icp(o_src, o_tgt, 20.0, 100)
    T = np.array(r.transform[:], np.float32).reshape(4, 4)
    print("gpu iterations", r.iterations, "state", r.convergence_state, "fitness", r.fitness, "| oracle", o["iterations"], o["state"], o["fitness"],
          "| dT", float(np.max(np.abs(T[:3, 3] - o["transform"][:3, 3]))), "dR", float(_rot_err(T, o["transform"])))
    assert r.converged == int(o["converged"]) == 1
    assert abs(r.iterations - o["iterations"]) <= 1                                  # a stop criterion sitting on its threshold may flip
    assert np.max(np.abs(T[:3, 3] - o["transform"][:3, 3])) < 1e-4 and _rot_err(T, o["transform"]) < 1e-5      # north-star tolerance (measured: 6e-8 m, 0 rad)
    assert abs(r.fitness - o["fitness"]) < 1e-3 * max(o["fitness"], 1e-3)
    # the correction undoes the drift: applied to the drifted pose it lands on the true one
    from bench import pose_to_T, T_to_pose
    fixed = T_to_pose(T.astype(np.float64) @ pose_to_T(revisit[-1][1].astype(np.float64)))
    assert np.linalg.norm(fixed[3:] - [4.0, 0, 0]) < 0.05 and abs(fixed[2]) < 2e-3
    assert r.fitness < 0.3                                                           # historyKeyframeFitnessScore gate (:665)
    p6 = np.array(r.pose6[:])
    assert np.allclose(p6, T_to_pose(T.astype(np.float64)), atol=1e-5)
    c.close()


def test_icp_reference_call_form_and_guards(oracle, synth, revisit):
    """performSCLoopClosure's own call form (base_key = 0: every cloud transformed by keyframe 0's pose) and its size guards."""
    import liorf_b200
    import pyicp
    c = liorf_b200.Context()
    for cl, p in revisit:
        c.addKeyframeCloud(cl, p)
    cur, pre = len(revisit) - 1, 4
    r = c.loopClosureICP(cur, pre, history_search_num=0, loop_index=0, icp_leaf=0.5, max_corr_dist=20.0, max_iters=100)
    clouds, poses = [k[0] for k in revisit], [k[1] for k in revisit]
    o_src = pyicp.loop_find_near_keyframes(clouds, poses, cur, 0, 0, 0.5, oracle)
    o_tgt = pyicp.loop_find_near_keyframes(clouds, poses, pre, 0, 0, 0.5, oracle)
    assert r.ran == 1 and r.n_source == len(o_src) and r.n_target == len(o_tgt)
    o = pyicp.icp(o_src, o_tgt, 20.0, 100)
    T = np.array(r.transform[:], np.float32).reshape(4, 4)
    assert r.converged == 1 and abs(r.iterations - o["iterations"]) <= 1
    assert np.max(np.abs(T[:3, 3] - o["transform"][:3, 3])) < 1e-4 and _rot_err(T, o["transform"]) < 1e-5      # north-star tolerance (measured: 6e-8 m, 0 rad)
    # both clouds are in their own sensor frames here, recorded at the same place: the correction is ~identity
    assert np.linalg.norm(T[:3, 3]) < 0.1 and _rot_err(T, np.eye(4, dtype=np.float32)) < 5e-3
    # guards (:655-656): a tiny cloud on either side → the ICP does not run
    tiny = liorf_b200.Context()
    tiny.addKeyframeCloud(revisit[0][0][:200], revisit[0][1]); tiny.addKeyframeCloud(revisit[1][0], revisit[1][1])
    assert tiny.loopClosureICP(0, 1, 0, 0).ran == 0 and tiny.loopClosureICP(1, 0, 0, 0).ran == 0
    tiny.close(); c.close()

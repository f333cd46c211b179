"""GPU parity: imageProjection deskew (a1/a2) and ScanContext (a10-a14) through the C ABI against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KITTI = dict(lidarMinRange=1.0, lidarMaxRange=1000.0, N_SCAN=64, downsampleRate=2, point_filter_num=5)


def _params_kw(d):
    return dict(N_SCAN=d["N_SCAN"], downsampleRate=d["downsampleRate"], point_filter_num=d["point_filter_num"],
                lidarMinRange=d["lidarMinRange"], lidarMaxRange=d["lidarMaxRange"])


@pytest.mark.parametrize("cfg", [KITTI, dict(KITTI, downsampleRate=1, point_filter_num=1), dict(KITTI, lidarMinRange=5.0, lidarMaxRange=60.0, point_filter_num=3)])
def test_project_point_cloud_deskew(oracle, synth, cfg):
    import liorf_b200
    c = liorf_b200.Context(**_params_kw(cfg))
    omega = (0.05, -0.08, 1.0)                                    # 1 rad/s yaw sweep: deskew is non-trivial
    raw = synth.scan(synth.HDL64, (0, 0, 0.2, 3, 1, 0), omega=omega, seed=7)
    t0 = 1000.0
    t_end = t0 + float(raw["time"][-1])
    imu_t, imu_rot, ptr = synth.imu_table(t0, t_end, omega, rate_hz=100.0, gyro_noise=1.56e-3, seed=3)
    out, n, kept = c.projectPointCloud(raw, t0, imu_t, imu_rot, ptr, True, want_kept_index=True)
    o_out, o_kept = oracle.project_point_cloud(raw, cfg, t0, imu_t, imu_rot, ptr, True)
    assert n == len(o_out) and np.array_equal(kept, o_kept)       # filters + order-preserving append: bit-exact
    assert np.array_equal(out[:, 3], o_out[:, 3])
    # rotation built from fp64-rounded vs glibc float sin/cos: <= few ulp of the matrix entries
    assert np.max(np.abs(out[:, :3] - o_out[:, :3])) < 5e-5
    moved = np.linalg.norm(out[:, :3] - synth.raw_to_xyzi(raw)[kept][:, :3], axis=1)
    assert moved.max() > 0.5                                      # the test really exercises the deskew
    # passthrough branch (:538)
    out2, n2 = c.projectPointCloud(raw, t0, deskew_enabled=False)
    o_out2, _ = oracle.project_point_cloud(raw, cfg, t0, imu_t, imu_rot, ptr, False)
    assert n2 == len(o_out2) and np.array_equal(out2, o_out2)
    c.close()


def test_project_point_cloud_edges(oracle, synth):
    import liorf_b200
    c = liorf_b200.Context()
    raw = synth.scan(synth.HDL64, (0, 0, 0, 0, 0, 0), seed=9)[:5000]
    # empty input
    out, n = c.projectPointCloud(raw[:0], 0.0, deskew_enabled=False)
    assert n == 0
    # everything filtered out (range window excludes all)
    c2 = liorf_b200.Context(lidarMinRange=500.0, lidarMaxRange=600.0)
    out, n = c2.projectPointCloud(raw, 0.0, deskew_enabled=False)
    assert n == 0
    # IMU table with a single usable row pair and points before/after the table (no extrapolation, :505-509)
    imu_t = np.array([10.02, 10.05], np.float64); imu_rot = np.array([[0, 0, 0], [0.01, -0.02, 0.03]], np.float64)
    out, n, kept = c.projectPointCloud(raw, 10.0, imu_t, imu_rot, 1, True, want_kept_index=True)
    o_out, o_kept = oracle.project_point_cloud(raw, KITTI, 10.0, imu_t, imu_rot, 1, True)
    assert np.array_equal(kept, o_kept) and np.max(np.abs(out - o_out)) < 5e-5
    c.close(); c2.close()


def test_project_then_downsample_chain(oracle, synth):
    """deskew output stays on the device and feeds downsampleCurrentScan without a host copy."""
    import liorf_b200
    c = liorf_b200.Context(downsampleRate=1, point_filter_num=1)
    raw = synth.scan(synth.HDL64, (0, 0, 0, 0, 0, 0), seed=11)
    out, n = c.projectPointCloud(raw, 0.0, deskew_enabled=False)
    ds, nds = c.downsampleCurrentScan(len(raw))
    o_ds, _, _ = oracle.voxel_grid(out, 0.4)
    assert nds == len(o_ds) and np.array_equal(ds, o_ds)
    c.close()


# ------------------------------------------------------------------------------------------------ ScanContext
def test_sc_make(ctx, oracle, synth):
    for k, pose in enumerate([(0, 0, 0, 0, 0, 0), (0, 0, 1.0, 40, 3, 0), (0, 0, -2.0, 100, -2, 0)]):
        cloud = synth.raw_to_xyzi(synth.scan(synth.HDL64, pose, seed=20 + k))
        ctx.makeAndSaveScancontextAndKeys(cloud)
        d, rk, sk = ctx.scGet(k)
        od, ork, osk = oracle.sc_make(cloud)
        assert np.array_equal(d, od)                               # polar bins: bit-exact
        assert np.array_equal(rk, ork) and np.array_equal(sk, osk)
    # empty cloud → all-zero descriptor
    ctx.makeAndSaveScancontextAndKeys(np.zeros((0, 4), np.float32))
    d, rk, sk = ctx.scGet(3)
    assert not d.any() and not rk.any()


def test_sc_make_from_device_scan(ctx, oracle, synth):
    raw = synth.scan(synth.HDL64, (0, 0, 0.4, 10, 0, 0), seed=31)
    out, n = ctx.projectPointCloud(raw, 0.0, deskew_enabled=False)
    ctx.makeAndSaveScancontextAndKeys()                            # SINGLE_SCAN_FULL: the full deskewed cloud (:1587-1595)
    d, rk, _ = ctx.scGet(0)
    od, ork, _ = oracle.sc_make(out)
    assert np.array_equal(d, od) and np.array_equal(rk, ork)


def test_sc_detect_loop_closure_sequence(ctx, oracle, synth):
    """reference semantics incl. the 31-entry early-out and the stale tree rebuilt every 10th call."""
    db = synth.sc_descriptors(90, seed=5)
    # make later entries revisit earlier places (column-shifted, noisy copies)
    q, src, shift = synth.sc_queries(db[:40], 50, seed=6)
    seq = np.concatenate([db[:40], q], 0)
    sc = oracle.SCManager()
    n_loops = 0
    for i in range(len(seq)):
        ctx.scAddDescriptors(seq[i:i + 1]); sc.save_descriptor(seq[i])
        g = ctx.detectLoopClosureID(); o = sc.detect()
        assert g[0] == o[0], (i, g, o)                             # loop id (or -1)
        assert g[1] == o[1]                                        # yaw = shift * 6 deg in rad, float
        if i >= 30:
            assert np.array_equal(g[3], o[3])                      # the three kNN candidates, in order
            assert abs(g[2] - o[2]) < 1e-12
        n_loops += g[0] >= 0
    assert n_loops > 5


def test_sc_query_batch(ctx, oracle, synth):
    K, Q = 6000, 300
    db = synth.sc_descriptors(K, seed=8)
    db[123] = db[77]                                               # exact duplicate keys → tie on (dist, idx)
    q, src, shift = synth.sc_queries(db, Q, seed=9)
    ctx.scAddDescriptors(db)
    loop, sh, dist, cand = ctx.scQueryBatch(q)
    keys = np.stack([oracle.sc_keys_from_desc(d)[0] for d in db])
    qkeys = np.stack([oracle.sc_keys_from_desc(d)[0] for d in q])
    o_loop, o_sh, o_dist, o_cand = oracle.sc_query_batch(keys, db, qkeys, q)
    assert np.array_equal(cand, o_cand)                            # candidate IDs bit-exact
    assert np.array_equal(loop, o_loop) and np.array_equal(sh, o_sh)   # loop ids and shifts bit-exact
    assert np.allclose(dist, o_dist, rtol=0, atol=1e-12, equal_nan=True)
    true = src >= 0
    assert (loop[true] == src[true]).mean() > 0.9                  # the planted loops are found
    assert np.array_equal(sh[true][loop[true] == src[true]], shift[true][loop[true] == src[true]])

"""Parity on the other BASELINE.json configurations (they are parity-test cases, not bench lines):
   config 3  Ouster OS1-128 shape (131k rays, scan leaf 0.2 m, map leaf 0.5 m)  — dense-map registration
   config 4  Livox non-repetitive pattern, 6-axis IMU at 200 Hz, point_filter_num 3, leaves 0.15 / 0.3
Everything goes through the C ABI and is compared with the oracle on the same seeded inputs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _make_case(synth, oracle, sensor, n_kf, scan_leaf, seed0):
    kfs = []
    for k in range(n_kf):
        p = np.array([0, 0, 0, 1.0 * k, 0, 0], np.float64)
        ds, _, _ = oracle.voxel_grid(synth.raw_to_xyzi(synth.scan(sensor, p, seed=seed0 + k)), scan_leaf)
        kfs.append((ds, p.astype(np.float32)))
    qp = np.array([0, 0, 0, 1.0 * (n_kf - 1), 0, 0], np.float64)
    raw = synth.scan(sensor, qp, seed=seed0 + 100)
    init = (qp + np.array([np.deg2rad(0.4), np.deg2rad(-0.3), np.deg2rad(1.0), 0.25, -0.1, 0.02])).astype(np.float32)
    return kfs, raw, init, qp.astype(np.float32)


def test_os1_128_dense_registration(oracle, synth):
    import liorf_b200
    kfs, raw, init, truth = _make_case(synth, oracle, synth.OS1_128, 8, 0.2, 300)
    c = liorf_b200.Context(N_SCAN=128, downsampleRate=1, point_filter_num=1, mappingSurfLeafSize=0.2, surroundingKeyframeMapLeafSize=0.5)
    for cl, p in kfs:
        c.addKeyframeCloud(cl, p)
    m = c.extractSurroundingKeyFrames(list(range(len(kfs))))
    o_map, _, _ = oracle.voxel_grid(np.concatenate([oracle.transform_cloud(cl, p) for cl, p in kfs]), 0.5)
    assert m == len(o_map) and np.array_equal(c.getLocalMap(), o_map)
    out, n = c.projectPointCloud(raw, 0.0, deskew_enabled=False)
    o_out, _ = oracle.project_point_cloud(raw, dict(lidarMinRange=1.0, lidarMaxRange=1000.0, N_SCAN=128, downsampleRate=1, point_filter_num=1), 0.0, None, np.zeros((1, 3)), 0, False)
    assert np.array_equal(out, o_out)
    ds, nds = c.downsampleCurrentScan(len(raw))
    o_ds, _, _ = oracle.voxel_grid(o_out, 0.2)
    assert nds == len(o_ds) > 20000 and np.array_equal(ds, o_ds)          # > 18.9k queries: the solver runs more than one round per CTA
    pose, tr = c.scan2MapOptimization(init, 30, force_all_iters=True)
    o = oracle.scan2map(o_ds, o_map, init, 30, force_all=True)
    gp = tr.poses()
    assert np.max(np.abs(gp[:, 3:] - o["trace"][:, 3:])) < 1e-4 and np.max(np.abs(gp[:, :3] - o["trace"][:, :3])) < 1e-5
    assert np.linalg.norm(pose[3:] - truth[3:]) < 0.05
    c.close()


def test_livox_deskew_200hz_and_registration(oracle, synth):
    import liorf_b200
    filt = dict(lidarMinRange=1.0, lidarMaxRange=1000.0, N_SCAN=6, downsampleRate=1, point_filter_num=3)
    c = liorf_b200.Context(N_SCAN=6, downsampleRate=1, point_filter_num=3, mappingSurfLeafSize=0.15, surroundingKeyframeMapLeafSize=0.3)
    omega = (0.02, -0.03, 1.0)                                             # 1 rad/s yaw sweep
    pose = np.array([0, 0, 0.3, 4.0, 0.5, 0], np.float64)
    raw = synth.scan(synth.LIVOX, pose, omega=omega, seed=400)
    t0 = 20.0
    it, rot, ptr = synth.imu_table(t0, t0 + float(raw["time"][-1]), omega, rate_hz=200.0, gyro_noise=1e-3, seed=4)
    assert ptr >= 20                                                       # ~22 rows per scan at 200 Hz
    out, n, kept = c.projectPointCloud(raw, t0, it, rot, ptr, True, want_kept_index=True)
    o_out, o_kept = oracle.project_point_cloud(raw, filt, t0, it, rot, ptr, True)
    assert np.array_equal(kept, o_kept) and np.all(kept % 3 == 0)
    assert np.max(np.abs(out - o_out)) < 5e-5
    ds, nds = c.downsampleCurrentScan(len(raw))
    o_ds, _, _ = oracle.voxel_grid(out, 0.15)                               # same input bytes → bit-exact
    assert nds == len(o_ds) and np.array_equal(ds, o_ds)
    # registration of an undistorted Livox scan against a small Livox map (forward-looking FoV: weakly constrained
    # sideways → exercises the degeneracy path on both sides)
    kfs, raw2, init, truth = _make_case(synth, oracle, synth.LIVOX, 10, 0.15, 500)
    for cl, p in kfs:
        c.addKeyframeCloud(cl, p)
    m = c.extractSurroundingKeyFrames(list(range(len(kfs))))
    o_map, _, _ = oracle.voxel_grid(np.concatenate([oracle.transform_cloud(cl, p) for cl, p in kfs]), 0.3)
    assert m == len(o_map)
    c.setCurrentScan(synth.raw_to_xyzi(raw2))
    ds2, n2 = c.downsampleCurrentScan(len(raw2))
    o_ds2, _, _ = oracle.voxel_grid(synth.raw_to_xyzi(raw2), 0.15)
    assert np.array_equal(ds2, o_ds2)
    pose_g, tr = c.scan2MapOptimization(init, 30, force_all_iters=True)
    o = oracle.scan2map(o_ds2, o_map, init, 30, force_all=True)
    assert tr.degenerate == int(o["state"][0])
    gp = tr.poses()
    assert np.max(np.abs(gp[:, 3:] - o["trace"][:, 3:])) < 1e-4 and np.max(np.abs(gp[:, :3] - o["trace"][:, :3])) < 1e-5
    c.close()


def test_degenerate_scene_projector(oracle):
    """ground plane only: x/y/yaw unobservable → eigenvalues < 100 → isDegenerate, X projected by matP (:1242-1271).
    The certificate must fail here and the exact cv::eigen path must run."""
    import liorf_b200
    rng = np.random.default_rng(7)
    g = np.zeros((60000, 4), np.float32); g[:, :2] = rng.uniform(-25, 25, size=(60000, 2)); g[:, 2] = -1.7 + rng.normal(scale=0.01, size=60000)
    c = liorf_b200.Context()
    m_ds = c.voxelGrid(g, 0.5)[0]
    c.setLocalMap(m_ds)
    q = np.zeros((20000, 4), np.float32); q[:, :2] = rng.uniform(-20, 20, size=(20000, 2)); q[:, 2] = -1.7 + rng.normal(scale=0.01, size=20000)
    c.setCurrentScan(q)
    ds, n = c.downsampleCurrentScan(len(q))
    init = np.array([0.01, -0.008, 0.02, 0.2, -0.1, 0.05], np.float32)
    pose, tr = c.scan2MapOptimization(init, 30, force_all_iters=True)
    o = oracle.scan2map(ds, m_ds, init, 30, force_all=True)
    assert tr.degenerate == 1 == int(o["state"][0])
    deg, P = c.getLMState()
    assert deg and np.allclose(P, o["state"][1:].reshape(6, 6), atol=2e-3)
    gp = tr.poses()
    assert np.max(np.abs(gp[:, 3:] - o["trace"][:, 3:])) < 1e-4 and np.max(np.abs(gp[:, :3] - o["trace"][:, :3])) < 1e-5
    assert abs(pose[5]) < 0.01 and abs(pose[0]) < 1e-3 and abs(pose[1]) < 1e-3      # observable dof converge; x,y,yaw stay near the guess
    assert abs(pose[3] - 0.2) < 0.05 and abs(pose[4] + 0.1) < 0.05
    c.close()


def test_global_map_filters(oracle, synth):
    """§8f-4: publishGlobalMap (src/mapOptmization.cpp:453-502) and saveMapService's map (:379-432) on the resident keyframes."""
    import liorf_b200
    kfs, _, _, _ = _make_case(synth, oracle, synth.HDL64, 14, 0.4, 900)
    for k, (cl, p) in enumerate(kfs):                                # spread them: 6 m apart so that the pose thinning (10 m) bites
        p[3] = 6.0 * k
    c = liorf_b200.Context()
    for cl, p in kfs:
        c.addKeyframeCloud(cl, p)
    # saveMapService: every keyframe, own pose, VoxelGrid(resolution)
    g = c.buildGlobalMap(search_radius=0.0, pose_density=0.0, leaf=0.8)
    o_all = np.concatenate([oracle.transform_cloud(cl, p) for cl, p in kfs])
    assert np.array_equal(g, oracle.voxel_grid(o_all, 0.8)[0])
    raw = c.buildGlobalMap(search_radius=0.0, pose_density=0.0, leaf=0.0)          # req.resolution == 0: no down-sampling
    assert np.array_equal(raw, o_all)
    # publishGlobalMap: radius 40 m around the newest pose, poses thinned to 10 m voxels, nearest-1 id recovery
    P = np.array([p for _, p in kfs], np.float32)[:, 3:6]
    d = ((P[-1] - P) ** 2).astype(np.float32).sum(1)
    near = [i for i in np.lexsort((np.arange(len(P)), d)) if d[i] < 40.0 ** 2]
    cent, _, _ = oracle.voxel_grid(np.concatenate([P[near], np.zeros((len(near), 1), np.float32)], 1), 10.0)
    ids = [int(np.argmin(((cc[:3] - P) ** 2).sum(1))) for cc in cent if not np.sqrt(((cc[:3] - P[-1]) ** 2).sum()) > 40.0]
    assert 2 <= len(ids) < len(near)
    o_sel = np.concatenate([oracle.transform_cloud(kfs[i][0], kfs[i][1]) for i in ids])
    g2 = c.buildGlobalMap(search_radius=40.0, pose_density=10.0, leaf=1.0)
    assert np.array_equal(g2, oracle.voxel_grid(o_sel, 1.0)[0])
    c.close()


def test_cloud_deskewed_wire_format(ctx, oracle, synth):
    """mapOptimization receives laserCloudSurfLast as cloud_info.cloud_deskewed (msg/cloud_info.msg:27): a PointCloud2 block of PCL
    PointXYZI points — 32 bytes each, x y z at 0, intensity at 16, padding elsewhere (include/utility.h:61).  Feeding that block as it
    is gives the same downsampled cloud, bit for bit, as feeding packed 16-byte points; an odd 21-byte layout exercises the unaligned path."""
    scan = synth.raw_to_xyzi(synth.scan(synth.HDL64, (0, 0, 0, 2.0, 0, 0), seed=synth.SEED0 + 5))[::7].copy()
    n = len(scan)
    ctx.setCurrentScan(scan)
    ref_ds, n_ref = ctx.downsampleCurrentScan(n)
    o_ds, _, _ = oracle.voxel_grid(scan, 0.4)
    assert n_ref == len(o_ds) and np.array_equal(ref_ds[:n_ref], o_ds)
    rng = np.random.default_rng(3)
    pcl = rng.integers(0, 255, size=(n, 32), dtype=np.uint8)                  # padding bytes are garbage on purpose
    pcl[:, 0:12] = scan[:, :3].copy().view(np.uint8).reshape(n, 12)
    pcl[:, 16:20] = scan[:, 3:4].copy().view(np.uint8).reshape(n, 4)
    ctx.setCurrentScanStrided(pcl, n, 32, 0, 16)
    ds, n_ds = ctx.downsampleCurrentScan(n)
    assert n_ds == n_ref and np.array_equal(ds[:n_ds], ref_ds[:n_ref])
    odd = rng.integers(0, 255, size=(n, 21), dtype=np.uint8)
    odd[:, 5:17] = scan[:, :3].copy().view(np.uint8).reshape(n, 12)
    odd[:, 1:5] = scan[:, 3:4].copy().view(np.uint8).reshape(n, 4)
    ctx.setCurrentScanStrided(odd, n, 21, 5, 1)
    ds2, n_ds2 = ctx.downsampleCurrentScan(n)
    assert n_ds2 == n_ref and np.array_equal(ds2[:n_ds2], ref_ds[:n_ref])

"""The oracle's restatement of the two ROS nodes, pinned to the reference's OWN source: /root/reference/src/mapOptmization.cpp and
src/imageProjection.cpp compiled UNCHANGED against the header stand-ins of oracle/shim_ros (recipe: oracle/Makefile → oracle/_ref/
libliorf_ref_mapopt.so, libliorf_ref_imageproj.so; wrappers oracle/ref_mapopt.cpp, oracle/ref_imageproj.cpp).

What is the reference's own code in these comparisons: control flow, indexing, thresholds, float / double mixes, member state that persists
across iterations and frames, queueing and call order of laserCloudInfoHandler (:236-275), updateInitialGuess (:899-958), extractNearby /
extractCloud (:975-1044), downsampleCurrentScan (:1061-1067), surfOptimization / combineOptimizationCoeffs / LMOptimization /
scan2MapOptimization / transformUpdate (:1074-1353), saveFrame / saveKeyFramesAndFactor (:1365-1384, 1503-1609), and of cloudHandler →
cachePointCloud / deskewInfo / imuDeskewInfo / projectPointCloud / deskewPoint / findRotation (src/imageProjection.cpp:191-598).
What is NOT: the arithmetic inside the third-party calls — the stand-ins forward VoxelGrid, ColPivHouseholderQR, cv::solve / eigen / inv,
getTransformation, Affine3f and tf quaternions to oracle/liorf_oracle.hpp, the kd-tree is the reference's vendored nanoflann instead of FLANN,
iSAM2 returns a new pose's initial value.  So a green test says: GIVEN those kernels, the oracle (and the library's host logic, where it is
used below) does exactly what the reference does — bit for bit.  CPU only; skipped when oracle/_ref was never built."""
import ctypes as C

import numpy as np
import pytest


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def refnodes(oracle):
    if not (oracle.RefMapOpt.available() and oracle.RefImageProjection.available()):
        pytest.skip("oracle/_ref/libliorf_ref_mapopt.so / libliorf_ref_imageproj.so were never built (needs /root/reference at build time)")
    return oracle


def _map_and_scan(oracle, case, map_leaf=0.5, scan_leaf=0.4):
    mp, _, _ = oracle.voxel_grid(np.concatenate([oracle.transform_cloud(c, p) for c, p in case["keyframes"]]), map_leaf)
    ds, _, _ = oracle.voxel_grid(case["scan"], scan_leaf)
    return ds, mp


def test_surf_and_lm_optimization_bit_exact_vs_reference(refnodes, kitti_case):
    """surfOptimization's flags and coefficients at the start pose, then 30 × {surfOptimization, combineOptimizationCoeffs, LMOptimization(i)} run by
    the reference's own member functions: every per-iteration pose, every correspondence count, isDegenerate and matP equal the oracle's, bit for bit."""
    o = refnodes
    ds, mp = _map_and_scan(o, kitti_case)
    R = o.RefMapOpt()
    R.set_scan_and_map(ds, mp)
    R.set_transform(kitti_case["init"])
    so = o.surf_optimization(ds, mp, kitti_case["init"])
    coeff, flag = R.surf_optimization()
    assert np.array_equal(flag, so["flag"]) and flag.sum() > 0.7 * len(ds)
    assert np.array_equal(_bits(coeff[flag == 1]), _bits(so["coeff"][flag == 1]))
    res = o.scan2map(ds, mp, kitti_case["init"], 30, True, use_ref_kdtree=True)           # same kd-tree (the reference's nanoflann) on both sides
    for it in range(30):
        conv, tf, nsel = R.iteration(it)
        assert np.array_equal(_bits(tf), _bits(res["trace"][it])), it
        assert nsel == res["nsel"][it]
    deg, P = R.lm_state()
    assert deg == int(res["state"][0]) and np.array_equal(_bits(P), _bits(res["state"][1:]))
    R.close()


def test_degenerate_projector_and_state_persistence_vs_reference(refnodes):
    """ground plane only: x / y / yaw unobservable → cv::eigen's small eigenvalues → isDegenerate, matP (:1242-1271); then a SECOND solve whose iteration 0
    has fewer than 50 correspondences keeps the first solve's projector (members persist across frames, SURVEY trap 9)."""
    o = refnodes
    rng = np.random.default_rng(7)
    g = np.zeros((20000, 4), np.float32); g[:, :2] = rng.uniform(-25, 25, size=(20000, 2)); g[:, 2] = -1.7 + rng.normal(scale=0.01, size=20000)
    mp, _, _ = o.voxel_grid(g, 0.5)
    q = np.zeros((6000, 4), np.float32); q[:, :2] = rng.uniform(-20, 20, size=(6000, 2)); q[:, 2] = -1.7 + rng.normal(scale=0.01, size=6000)
    ds, _, _ = o.voxel_grid(q, 0.4)
    init = np.array([0.01, -0.008, 0.02, 0.2, -0.1, 0.05], np.float32)
    R = o.RefMapOpt()
    R.set_scan_and_map(ds, mp)
    R.set_transform(init)
    res = o.scan2map(ds, mp, init, 12, True, use_ref_kdtree=True)
    for it in range(12):
        conv, tf, nsel = R.iteration(it)
        assert np.array_equal(_bits(tf), _bits(res["trace"][it])), it
    deg, P = R.lm_state()
    assert deg == 1 == int(res["state"][0]) and np.array_equal(_bits(P), _bits(res["state"][1:]))
    # second frame: 40 scan points only → N_sel < 50 at every iteration: LMOptimization returns false, pose untouched, projector kept
    ds2 = ds[:40]
    R.set_scan_and_map(ds2, mp)
    R.set_transform(init)
    res2 = o.scan2map(ds2, mp, init, 3, True, res["state"], use_ref_kdtree=True)
    for it in range(3):
        conv, tf, nsel = R.iteration(it)
        assert not conv and nsel < 50 and np.array_equal(_bits(tf), _bits(init))
    deg2, P2 = R.lm_state()
    assert deg2 == 1 and np.array_equal(_bits(P2), _bits(P)) and np.array_equal(_bits(res2["state"]), _bits(res["state"]))
    R.close()


@pytest.mark.parametrize("imu_type", [0, 1])
def test_scan2map_optimization_whole_function_vs_reference(refnodes, kitti_case, imu_type):
    """scan2MapOptimization() as one call (:1295-1321): guards, kd-tree build, the loop with the reference's convergence break, transformUpdate (9-axis
    roll / pitch slerp for imuType 1, clamps) — the final transformTobeMapped equals oracle scan2map (early exit) + oracle transform_update."""
    o = refnodes
    ds, mp = _map_and_scan(o, kitti_case)
    rpy = np.array([0.004, -0.003, 0.1], np.float32)
    R = o.RefMapOpt(imuType=imu_type, imuRPYWeight=0.05, rotation_tollerance=0.5, z_tollerance=20.0)
    R.set_scan_and_map(ds, mp)
    got = R.scan2map(kitti_case["init"], imu_available=1, rpy_init=rpy)
    res = o.scan2map(ds, mp, kitti_case["init"], 30, False, use_ref_kdtree=True)
    assert 2 <= res["iters"] < 30
    want = o.transform_update(res["tf"], 1, imu_type, float(rpy[0]), float(rpy[1]), 0.05, 0.5, 20.0)
    assert np.array_equal(_bits(got), _bits(want))
    # the guard of :1300: 30 points or fewer → nothing runs, the pose is untouched
    R.set_scan_and_map(ds[:30], mp)
    assert np.array_equal(_bits(R.scan2map(kitti_case["init"])), _bits(kitti_case["init"]))
    R.close()


def _host(lib):
    vp = lambda a: a.ctypes.data_as(C.c_void_p)

    def extract_nearby(poses, times, t_cur, radius, density):
        P = np.ascontiguousarray(np.array(poses, np.float32)); T = np.ascontiguousarray(np.array(times, np.float64))
        ids = np.zeros(4096, np.int32); n = C.c_int(0)
        assert lib.liorf_host_extract_nearby(vp(P), vp(T), len(P), C.c_double(t_cur), C.c_float(radius), C.c_float(density), vp(ids), 4096, C.byref(n)) == 0
        return ids[:n.value].copy()

    def save_frame(last, cur, d, a):
        last = None if last is None else np.ascontiguousarray(last, np.float32)
        return lib.liorf_host_save_frame(None if last is None else vp(last), vp(cur), C.c_float(d), C.c_float(a)) == 1
    return vp, extract_nearby, save_frame


@pytest.mark.parametrize("imu_type,heading", [(0, 0), (1, 1)])
def test_sequence_through_the_reference_node(refnodes, imu_type, heading):
    """A 40-frame drive, one liorf::cloud_info per scan through the reference's laserCloudInfoHandler, against the same frames through the oracle's kernels
    glued by the LIBRARY's host logic (liorf_host_update_initial_guess / extract_nearby / transform_update / save_frame — host-only entries of the C ABI,
    no GPU): after every frame the pose is bit-equal, so are the keyframe decisions and both cloud sizes; at the end every stored keyframe cloud and pose,
    the last local map and the ScanContext loop answer are equal.  Odometry and IMU availability toggle along the way (all three branches of
    updateInitialGuess); the (1, 1) variant is a 9-axis IMU with heading initialisation (the slerp of transformUpdate)."""
    import liorf_b200
    from liorf_b200 import GuessState, CloudInfoGuess
    from bench_common import Sequence
    o = refnodes
    lib = liorf_b200.load_library()
    vp, extract_nearby, save_frame = _host(lib)
    N = 40
    seq = Sequence(N)
    R = o.RefMapOpt(imuType=imu_type, useImuHeadingInitialization=heading, imuRPYWeight=0.02)
    st = GuessState(); tf = np.zeros(6, np.float32); lm = np.zeros(37, np.float32)
    kf_clouds, kf_poses, kf_times = [], [], []
    sc = o.SCManager()
    rng = np.random.default_rng(5)
    mp = None
    for i in range(N):
        raw, (t0, it, rot, ptr) = seq.frame(i)
        cloud, _ = o.project_point_cloud(raw, seq.filters, t0, it, rot, ptr, True)
        odo = (seq.poses[i] + np.concatenate([rng.normal(scale=np.deg2rad(0.1), size=3), rng.normal(scale=0.02, size=3)])).astype(np.float32)
        rpy = (seq.poses[i][:3] + rng.normal(scale=1e-3, size=3)).astype(np.float32)
        guess = np.array([odo[3], odo[4], odo[5], odo[0], odo[1], odo[2]], np.float32)
        imu_av, odo_av = int(i % 7 != 6), int(i % 5 != 4 and i % 11 != 3)
        R.cloud_info(t0, cloud, imu_av, odo_av, rpy, guess)
        rs = R.state()
        ci = CloudInfoGuess(imu_av, odo_av, *[float(v) for v in rpy], *[float(v) for v in guess])
        assert lib.liorf_host_update_initial_guess(C.byref(st), int(not kf_clouds), C.byref(ci), heading, imu_type, vp(tf)) == 0
        m_ds = 0
        if kf_clouds:
            ids = extract_nearby(kf_poses, kf_times, t0, 50.0, 2.0)
            mp, _, _ = o.voxel_grid(np.concatenate([o.transform_cloud(kf_clouds[k], kf_poses[k]) for k in ids]), 0.5); m_ds = len(mp)
        ds, _, _ = o.voxel_grid(cloud, 0.4)
        if kf_clouds and len(ds) > 30:
            r = o.scan2map(ds, mp, tf, 30, False, lm, use_ref_kdtree=True)
            tf, lm = r["tf"], r["state"]
            lib.liorf_host_transform_update(vp(tf), imu_av, imu_type, C.c_float(rpy[0]), C.c_float(rpy[1]), C.c_float(0.02), C.c_float(1000.0), C.c_float(1000.0))
        if save_frame(kf_poses[-1] if kf_poses else None, tf, 1.0, 0.2):
            kf_clouds.append(ds); kf_poses.append(tf.copy()); kf_times.append(t0)
            sc.make_and_save(cloud)
        assert np.array_equal(_bits(tf), _bits(rs["tf"])), (i, tf, rs["tf"])
        assert rs["keyframes"] == len(kf_clouds) and rs["n_ds"] == len(ds) and rs["m_ds"] == m_ds and rs["sc_entries"] == len(kf_clouds), i
    assert 15 <= len(kf_clouds) < N
    for k in range(len(kf_clouds)):
        assert np.array_equal(_bits(R.get_cloud(100 + k)), _bits(kf_clouds[k]))
        p, t = R.keypose(k)
        assert np.array_equal(_bits(p), _bits(kf_poses[k])) and t == kf_times[k]
    assert np.array_equal(_bits(R.get_cloud(1)), _bits(mp))                        # laserCloudSurfFromMapDS of the last frame that built one
    for _ in range(3):                                                              # detectLoopClosureID of the node's SCManager (incl. the call counter)
        lid, yaw = R.sc_detect()
        olid, oyaw, _, _ = sc.detect()
        assert lid == olid and np.float32(yaw) == np.float32(oyaw)
    R.close()


def test_map_cache_is_result_neutral_in_the_reference(refnodes, kitti_case):
    """laserCloudMapContainer (:1022-1032) only memoises transformPointCloud: a node that saw the keyframes one by one (cache warm) and the oracle's
    from-scratch map build agree on the local map — the property the library's resident-map reuse relies on."""
    o = refnodes
    R = o.RefMapOpt()
    kfs = kitti_case["keyframes"]
    t = 100.0
    for cl, p in kfs:                                   # each keyframe enters through the handler: first frame = keyframe at its guess, later ones are solved
        R.cloud_info(t, cl, 0, 0, None, None)
        t += 0.1
    s = R.state()
    assert s["keyframes"] >= 1 and s["m_ds"] > 0
    poses = [R.keypose(k)[0] for k in range(s["keyframes"])]
    times = [R.keypose(k)[1] for k in range(s["keyframes"])]
    clouds = [R.get_cloud(100 + k) for k in range(s["keyframes"])]
    # rebuild the last frame's map from scratch with the oracle from what was stored BEFORE that frame added its own keyframe (if it did)
    import liorf_b200
    vp, extract_nearby, _ = _host(liorf_b200.load_library())
    t_last = t - 0.1
    n_before = s["keyframes"] - (1 if times[-1] == t_last else 0)
    ids = extract_nearby(poses[:n_before], times[:n_before], t_last, 50.0, 2.0)
    mp, _, _ = o.voxel_grid(np.concatenate([o.transform_cloud(clouds[k], poses[k]) for k in ids]), 0.5)
    assert np.array_equal(_bits(R.get_cloud(1)), _bits(mp))
    R.close()


# ------------------------------------------------------------------------------------------------------------ ImageProjection
def _imu_stream(rng, t0, t1, rate, omega):
    n = int((t1 - t0) * rate) + 1
    t = np.sort(t0 + np.arange(n) / rate + rng.uniform(-0.2 / rate, 0.2 / rate, n))
    return t, rng.normal(0, 0.3, (n, 3)) + np.asarray(omega)


@pytest.mark.parametrize("cfg", ["kitti", "livox"])
def test_image_projection_node_vs_oracle(refnodes, synth, cfg):
    """IMU samples through imuHandler, scans through cloudHandler, the published liorf/cloud_info read back: the deskewed cloud equals the oracle's
    imu_deskew_info + project_point_cloud bit for bit (range / ring / raw-index filters, first-kept-point reference, findRotation's fp64 interpolation), the
    node answers two scans late (:209-215), the IMU table of the library's host entry equals the oracle's.  KITTI filters (1/2 rings, every 5th point,
    100 Hz) and the Livox configuration (6 lines, every 3rd point, 200 Hz)."""
    import liorf_b200
    o = refnodes
    rng = np.random.default_rng(11)
    if cfg == "kitti":
        filt = dict(lidarMinRange=1.0, lidarMaxRange=1000.0, N_SCAN=64, downsampleRate=2, point_filter_num=5); sensor, rate = synth.HDL64, 100.0
    else:
        filt = dict(lidarMinRange=1.0, lidarMaxRange=1000.0, N_SCAN=6, downsampleRate=1, point_filter_num=3); sensor, rate = synth.LIVOX, 200.0
    omega = (0.02, -0.03, 0.8)
    R = o.RefImageProjection(imuRate=rate, **{k: filt[k] for k in ("N_SCAN", "downsampleRate", "point_filter_num", "lidarMinRange", "lidarMaxRange")})
    stamps, gyro = _imu_stream(rng, 999.5, 1001.5, rate, omega)
    for s, g in zip(stamps, gyro):
        R.imu(s, g)
    scans = [(1000.0 + 0.1 * k + 0.0037, synth.scan(sensor, np.array([0, 0, 0, 2.0 * k, 0, 0], np.float64), omega=omega, seed=77 + k)) for k in range(6)]
    published = 0
    for k, (cur, raw) in enumerate(scans):
        n = R.cloud(cur, raw)
        assert n == max(0, k - 1)                                                   # two scans are held back
        if n > published:
            published = n
            c0, r0 = scans[k - 2]
            info = R.last_info()
            end = c0 + float(r0["time"][-1])                                        # timeScanEnd: the LAST point's time (:283)
            od = o.imu_deskew_info(stamps, gyro, c0, end)
            g = liorf_b200.imuDeskewInfo(stamps, gyro, c0, end)
            assert info["stamp"] == c0 and info["imuAvailable"] == 1 and od["available"] and g["available"]
            assert g["imu_pointer_cur"] == od["imu_pointer_cur"] and np.array_equal(g["imu_rot"], od["imu_rot"]) and np.array_equal(g["imu_time"], od["imu_time"])
            cloud, kept = o.project_point_cloud(r0, filt, c0, od["imu_time"], od["imu_rot"], od["imu_pointer_cur"], True)
            assert len(cloud) > 1000 and info["cloud"].shape == cloud.shape
            assert np.array_equal(_bits(info["cloud"]), _bits(cloud))
            assert np.abs(cloud[:, :3] - synth.raw_to_xyzi(r0)[kept][:, :3]).max() > 1e-3          # the deskew did move points
    assert published == 4
    R.close()


def test_image_projection_gates_vs_oracle(refnodes, synth):
    """no "time" field → deskewFlag = -1 → points pass through unrotated (:313-326, :538); IMU stream that does not cover the scan → deskewInfo returns false and
    nothing is published (:336-340) — the gate the library's liorf_host_imu_deskew_info(check_gate = 1) reports as "not available"."""
    import liorf_b200
    o = refnodes
    rng = np.random.default_rng(12)
    filt = dict(lidarMinRange=1.0, lidarMaxRange=1000.0, N_SCAN=64, downsampleRate=2, point_filter_num=5)
    stamps, gyro = _imu_stream(rng, 999.5, 1001.0, 100.0, (0, 0, 0.5))
    raws = [synth.scan(synth.HDL64, np.array([0, 0, 0, 1.0 * k, 0, 0], np.float64), omega=(0, 0, 0.5), seed=5 + k) for k in range(3)]
    R = o.RefImageProjection()
    for s, g in zip(stamps, gyro):
        R.imu(s, g)
    for k in range(3):
        n = R.cloud(1000.0 + 0.1 * k, raws[k], has_time=False)
    assert n == 1
    info = R.last_info()
    cloud, kept = o.project_point_cloud(raws[0], filt, 1000.0, np.zeros(1), np.zeros((1, 3)), 0, False)
    assert np.array_equal(_bits(info["cloud"]), _bits(cloud)) and np.array_equal(_bits(cloud), _bits(synth.raw_to_xyzi(raws[0])[kept]))
    R.close()
    R = o.RefImageProjection()
    late = stamps[stamps > 1000.05]                                                 # the first IMU sample is younger than the scan start
    for s, g in zip(late, gyro[stamps > 1000.05]):
        R.imu(s, g)
    for k in range(3):
        n = R.cloud(1000.0 + 0.1 * k, raws[k])
    assert n == 0 and R.last_info() is None
    end = 1000.0 + float(raws[0]["time"][-1])
    assert not liorf_b200.imuDeskewInfo(late, gyro[stamps > 1000.05], 1000.0, end, check_gate=True)["available"]
    R.close()


# ------------------------------------------------------------------------------------------------------------ §8f-3 / §8f-4 next to the path
def _keyframe_set(oracle, synth, n, spacing):
    kfs = []
    for k in range(n):
        p = np.array([0.002 * k, -0.001 * k, 0.01 * k, spacing * k, 0.3 * np.sin(k), 0.02 * k], np.float32)
        cl, _, _ = oracle.voxel_grid(synth.raw_to_xyzi(synth.scan(synth.HDL64, p.astype(np.float64), seed=900 + k)), 0.4)
        kfs.append((cl, p))
    return kfs


def test_loop_find_near_keyframes_vs_reference(refnodes, synth):
    """the sub-map assembly of the loop-closure ICP (loopFindNearKeyframes :821-844, as performSCLoopClosure calls it with base_key = 0 and as
    performRSLoopClosure does with -1): oracle/pyicp.py's restatement — what tests/test_gpu_icp.py checks the library against — equals the reference's own."""
    import pyicp
    o = refnodes
    kfs = _keyframe_set(o, synth, 9, 1.5)
    R = o.RefMapOpt(loopClosureICPSurfLeafSize=0.3)
    for k, (cl, p) in enumerate(kfs):
        R.add_keyframe(cl, p, 10.0 + 0.5 * k)
    clouds = [c for c, _ in kfs]; poses = [p for _, p in kfs]
    for key, num, base in ((8, 0, 0), (3, 25, 0), (3, 2, -1), (0, 1, -1), (8, 3, 0)):
        want = pyicp.loop_find_near_keyframes(clouds, poses, key, num, base, 0.3, o)
        got = R.loop_find_near_keyframes(key, num, base)
        assert len(want) > 1000 and got.shape == want.shape and np.array_equal(_bits(got), _bits(want)), (key, num, base)
    R.close()


def test_publish_global_map_vs_reference(refnodes, synth):
    """publishGlobalMap (:453-502): radius search around the newest key pose, pose thinning by a VoxelGrid, nearest-1 id recovery, the distance gate on
    the voxel centroid, transform + concatenate + VoxelGrid — the numpy restatement tests/test_gpu_configs.py::test_global_map_filters holds the library
    to, here against the reference's own function (one subscriber on liorf/mapping/map_global)."""
    o = refnodes
    kfs = _keyframe_set(o, synth, 14, 6.0)
    R = o.RefMapOpt(globalMapVisualizationSearchRadius=40.0, globalMapVisualizationPoseDensity=10.0, globalMapVisualizationLeafSize=1.0)
    assert R.publish_global_map() is None                                          # no key poses yet: early return (:458)
    for k, (cl, p) in enumerate(kfs):
        R.add_keyframe(cl, p, 10.0 + 0.5 * k)
    P = np.array([p for _, p in kfs], np.float32)[:, 3:6]
    d = ((P[-1] - P) ** 2).astype(np.float32).sum(1)
    near = [i for i in np.lexsort((np.arange(len(P)), d)) if d[i] < 40.0 ** 2]
    cent, _, _ = o.voxel_grid(np.concatenate([P[near], np.zeros((len(near), 1), np.float32)], 1), 10.0)
    ids = [int(np.argmin(((cc[:3] - P) ** 2).sum(1))) for cc in cent if not np.sqrt(((cc[:3] - P[-1]) ** 2).sum()) > 40.0]
    assert 2 <= len(ids) < len(near)
    want = o.voxel_grid(np.concatenate([o.transform_cloud(kfs[i][0], kfs[i][1]) for i in ids]), 1.0)[0]
    got = R.publish_global_map()
    assert got.shape == want.shape and np.array_equal(_bits(got), _bits(want))
    R.close()


def test_headline_config_full_scale_vs_reference_and_committed_gpu_pose(refnodes):
    """BASELINE config 1 at full size (50 full-density keyframes, N_ds = 13 364, M = 56 460, 30 forced iterations): the reference's own member functions
    and the oracle agree on every iteration bit for bit — and the final pose is, bit for bit, the `final_pose` the B200 run of bench.py committed under
    profiles/ for the same input bytes (sha256 in both lines).  GPU == oracle == the reference's code, on the configuration the < 1 ms target is quoted on."""
    import json
    import os
    import bench
    o = refnodes
    inst = bench.make_single_inputs("kitti64_single")
    kfs = [o.voxel_grid(s, 0.4)[0] for s in inst["scans"]]
    mp, _, _ = o.voxel_grid(np.concatenate([o.transform_cloud(c, p.astype(np.float32)) for c, p in zip(kfs, inst["poses"])]), 0.5)
    ds, _, _ = o.voxel_grid(inst["scan"], 0.4)
    assert (len(ds), len(mp)) == (13364, 56460)
    res = o.scan2map(ds, mp, inst["init"], 30, True, use_ref_kdtree=True)
    R = o.RefMapOpt()
    R.set_scan_and_map(ds, mp)
    R.set_transform(inst["init"])
    for it in range(30):
        conv, tf, nsel = R.iteration(it)
        assert np.array_equal(_bits(tf), _bits(res["trace"][it])) and nsel == res["nsel"][it], it
    R.close()
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_bench_n1.json")
    if os.path.exists(path):
        line = [json.loads(l) for l in open(path) if l.strip().startswith("{")][-1]
        if line["config"].get("input_sha256") == inst["sha256"]:
            assert np.array_equal(_bits(np.array(line["final_pose"], np.float32)), _bits(tf))


@pytest.mark.parametrize("cfg", ["os1_128_dense", "livox_deskew"])
def test_other_configurations_vs_reference(refnodes, synth, cfg):
    """BASELINE configs 3 and 4 at the sizes of tests/test_gpu_configs.py (same generator, same seeds): the node configured as lio_sam_ouster.yaml with the
    0.2 m scan leaf BASELINE asks for / as lio_sam_livox.yaml (0.15 / 0.3 m) — downsampleCurrentScan by the reference's own function, then its iterations — against
    the oracle.  The Livox field of view is weakly constrained sideways, so this also walks the degenerate branch on real-looking data."""
    o = refnodes
    sensor, n_scan, scan_leaf, map_leaf, n_kf, seed0 = ((synth.OS1_128, 128, 0.2, 0.5, 8, 300) if cfg == "os1_128_dense" else (synth.LIVOX, 6, 0.15, 0.3, 10, 500))
    kfs = []
    for k in range(n_kf):
        p = np.array([0, 0, 0, 1.0 * k, 0, 0], np.float64)
        ds, _, _ = o.voxel_grid(synth.raw_to_xyzi(synth.scan(sensor, p, seed=seed0 + k)), scan_leaf)
        kfs.append((ds, p.astype(np.float32)))
    qp = np.array([0, 0, 0, 1.0 * (n_kf - 1), 0, 0], np.float64)
    scan = synth.raw_to_xyzi(synth.scan(sensor, qp, seed=seed0 + 100))
    init = (qp + np.array([np.deg2rad(0.4), np.deg2rad(-0.3), np.deg2rad(1.0), 0.25, -0.1, 0.02])).astype(np.float32)
    mp, _, _ = o.voxel_grid(np.concatenate([o.transform_cloud(c, p) for c, p in kfs]), map_leaf)
    R = o.RefMapOpt(N_SCAN=n_scan, Horizon_SCAN=1024 if cfg == "os1_128_dense" else 4000, mappingSurfLeafSize=scan_leaf, surroundingKeyframeMapLeafSize=map_leaf)
    R.set_map(mp)
    pose, iters, _ = R.bench_step(scan, init, 30, True)                     # downsampleCurrentScan + kd-tree + 30 forced iterations, all the reference's
    ds, _, _ = o.voxel_grid(scan, scan_leaf)
    assert R.state()["n_ds"] == len(ds) and np.array_equal(_bits(R.get_cloud(0)), _bits(ds))
    res = o.scan2map(ds, mp, init, 30, True, use_ref_kdtree=True)
    assert iters == 30 and np.array_equal(_bits(pose), _bits(res["tf"]))
    deg, P = R.lm_state()
    assert deg == int(res["state"][0]) and np.array_equal(_bits(P), _bits(res["state"][1:]))
    R.close()


def test_imu_rpy_init_sample_vs_reference(refnodes, synth):
    """9-axis IMU (imuType 1): cloudInfo.imu{Roll,Pitch,Yaw}Init come from the LAST queued sample at or before the scan start that survived the pop (:371-375).
    The library's liorf_host_imu_deskew_info reports that sample as rpy_index; its orientation through tf's getRPY (the oracle's restatement behind the stand-in) must be
    what the reference node publishes."""
    import liorf_b200
    o = refnodes
    rng = np.random.default_rng(13)
    stamps, gyro = _imu_stream(rng, 999.5, 1001.0, 100.0, (0.0, 0.0, 0.3))
    rpy = np.stack([0.02 * np.sin(stamps), 0.03 * np.cos(stamps), 0.3 * (stamps - 999.5)], 1)          # a smooth attitude, different at every sample
    quats = []
    for r, p, y in rpy:
        cy, sy, cp, sp, cr, sr = np.cos(y / 2), np.sin(y / 2), np.cos(p / 2), np.sin(p / 2), np.cos(r / 2), np.sin(r / 2)
        quats.append([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy])
    quats = np.array(quats, np.float64)
    R = o.RefImageProjection(imuType=1)                                    # extrinsicRPY absent → identity (the stand-in's Map of an empty vector)
    for s, g, q in zip(stamps, gyro, quats):
        R.imu(s, g, q)
    raws = [synth.scan(synth.HDL64, np.array([0, 0, 0, 1.0 * k, 0, 0], np.float64), omega=(0, 0, 0.3), seed=40 + k) for k in range(4)]
    curs = [1000.0 + 0.1 * k + 0.0042 for k in range(4)]
    seen = 0
    for k in range(4):
        n = R.cloud(curs[k], raws[k])
        if n > seen:
            seen = n
            c0 = curs[k - 2]; end = c0 + float(raws[k - 2]["time"][-1])
            info = R.last_info()
            g = liorf_b200.imuDeskewInfo(stamps, gyro, c0, end)
            i = g["rpy_index"]
            assert i >= 0 and stamps[i] <= c0 and (i + 1 == len(stamps) or stamps[i + 1] > c0)
            want = rpy[i]
            assert np.allclose(info["rpy_init"], want, atol=1e-6), (info["rpy_init"], want)
            assert not np.allclose(info["rpy_init"], rpy[i - 1], atol=1e-5)                          # the neighbouring sample would not have passed
    assert seen == 2
    R.close()


def test_keyframe_selection_at_scale_vs_reference(refnodes, synth):
    """extractNearby + extractCloud (:975-1044) on 700 key poses of a closed drive that revisits its first leg: every keyframe cloud is ONE point whose intensity is the
    keyframe id, so the reference's laserCloudSurfFromMap spells out its selection — ids, order, duplicates (radius set + time tail, SURVEY trap 10), the distance gate on the
    voxel centroid — and the library's liorf_host_extract_nearby must spell the same list, at several query times (all keyframes older than 10 s; a 10 s tail; in between)."""
    import liorf_b200
    o = refnodes
    vp, extract_nearby, _ = _host(liorf_b200.load_library())
    traj = synth.street_trajectory(1400, loop=True)[::2]                       # 700 poses, 1.6 m apart
    n = len(traj)
    poses = np.concatenate([traj[:, :3], traj[:, 3:6]], 1).astype(np.float32)
    times = 100.0 + 0.2 * np.arange(n)
    R = o.RefMapOpt()
    for k in range(n):
        R.add_keyframe(np.array([[0.5, 0.0, 0.0, float(k)]], np.float32), poses[k], times[k])
    checked = 0
    for t_cur in (times[-1] + 0.1, times[-1] + 4.0, times[-1] + 30.0):
        R.extract_surrounding_keyframes(t_cur)
        got = R.get_cloud(2)[:, 3].astype(np.int64)
        want = extract_nearby(poses, times, t_cur, 50.0, 2.0)
        assert len(want) >= 20 and np.array_equal(got, want), (t_cur, got[:12], want[:12])
        checked += 1
    # a shorter history whose newest pose sits on the revisited leg: old and new keyframes of the same street are selected together
    m = int(np.argmin(np.linalg.norm(poses[300:, 3:5] - poses[5, 3:5], axis=1))) + 300
    R2 = o.RefMapOpt()
    for k in range(m + 1):
        R2.add_keyframe(np.array([[0.5, 0.0, 0.0, float(k)]], np.float32), poses[k], times[k])
    R2.extract_surrounding_keyframes(times[m] + 0.05)
    got = R2.get_cloud(2)[:, 3].astype(np.int64)
    want = extract_nearby(poses[:m + 1], times[:m + 1], times[m] + 0.05, 50.0, 2.0)
    assert np.array_equal(got, want) and got.min() < 40 and got.max() == m
    R.close(); R2.close()
    assert checked == 3


def test_scalar_host_entries_fuzzed_against_the_reference_members(refnodes):
    """liorf_host_update_initial_guess / liorf_host_transform_update / liorf_host_save_frame against the reference's own updateInitialGuess (:899-958), transformUpdate
    (:1323-1353) and saveFrame (:1365-1384), called directly on the node: random sequences (availability flags toggling, large attitudes up to the gimbal region, 6- and
    9-axis, heading initialisation on / off), thousands of random pose pairs around the keyframe thresholds — bit for bit / decision for decision."""
    import liorf_b200
    from liorf_b200 import GuessState, CloudInfoGuess
    o = refnodes
    lib = liorf_b200.load_library()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rng = np.random.default_rng(21)
    for trial, (imu_type, heading) in enumerate([(0, 0), (1, 1), (1, 0), (0, 1)] * 3):
        R = o.RefMapOpt(imuType=imu_type, useImuHeadingInitialization=heading, imuRPYWeight=0.07, rotation_tollerance=0.9, z_tollerance=3.0)
        st = GuessState(); tf = np.zeros(6, np.float32); rtf = np.zeros(6, np.float32)
        scale = [0.05, 0.3, 1.2][trial % 3]                                  # attitude spread: small, large, towards +-pi/2 pitch
        odo = rng.normal(scale=[scale, scale, 1.0, 5, 5, 1], size=6)
        for k in range(60):
            odo = odo + rng.normal(scale=[0.01 * scale * 10, 0.01 * scale * 10, 0.05, 0.8, 0.3, 0.05], size=6)
            odo[1] = np.clip(odo[1], -1.5, 1.5)
            rpy = odo[:3] + rng.normal(scale=3e-3, size=3)
            ci11 = np.array([rng.random() < 0.8, rng.random() < 0.7, *rpy, odo[3], odo[4], odo[5], odo[0], odo[1], odo[2]], np.float32)
            have_kf = k > 0
            ci = CloudInfoGuess(int(ci11[0]), int(ci11[1]), *[float(v) for v in ci11[2:]])
            assert lib.liorf_host_update_initial_guess(C.byref(st), int(not have_kf), C.byref(ci), heading, imu_type, vp(tf)) == 0
            rtf = R.update_initial_guess(ci11, have_kf, rtf)
            assert np.array_equal(_bits(tf), _bits(rtf)), (trial, k, tf, rtf)
            # a solver step in between (the same perturbation on both sides), then transformUpdate
            step = rng.normal(scale=[2e-3, 2e-3, 5e-3, 0.05, 0.05, 0.02], size=6).astype(np.float32)
            tf = (tf + step).astype(np.float32); rtf = (rtf + step).astype(np.float32)
            lib.liorf_host_transform_update(vp(tf), int(ci11[0]), imu_type, C.c_float(ci11[2]), C.c_float(ci11[3]), C.c_float(0.07), C.c_float(0.9), C.c_float(3.0))
            rtf = R.transform_update(ci11, rtf)
            assert np.array_equal(_bits(tf), _bits(rtf)), (trial, k, "transformUpdate", tf, rtf)
        R.close()
    R = o.RefMapOpt(surroundingkeyframeAddingDistThreshold=1.0, surroundingkeyframeAddingAngleThreshold=0.2)
    n_true = 0
    for k in range(3000):
        last = rng.normal(scale=[0.3, 0.3, 2.0, 50, 50, 2], size=6).astype(np.float32)
        d = rng.normal(scale=[0.12, 0.12, 0.12, 0.6, 0.6, 0.3], size=6).astype(np.float32)
        cur = (last + d).astype(np.float32)
        want = R.save_frame(last, cur)
        got = lib.liorf_host_save_frame(vp(last), vp(cur), C.c_float(1.0), C.c_float(0.2)) == 1
        assert want == got, (k, last, cur)
        n_true += int(want)
    assert 300 < n_true < 2700                                                # both outcomes well represented
    assert R.save_frame(None, np.zeros(6, np.float32)) is True and lib.liorf_host_save_frame(None, vp(np.zeros(6, np.float32)), C.c_float(1.0), C.c_float(0.2)) == 1
    R.close()


def test_surf_optimization_branches_fuzzed_against_the_reference(refnodes):
    """surfOptimization (:1074-1143) on maps that are NOT friendly: noisy planes at several noise levels, sparse volumes (5th neighbour beyond 1 m → no fit), points far
    from the sensor origin and near it (the s weight's sqrt(sqrt(range)) denominator), so that every branch — distance gate, plane validity at 0.2, s > 0.1 — is taken often;
    flags and coefficient bits of every point equal the oracle's; the reference's kd-tree (nanoflann) neighbour ORDER is what the oracle's kd-tree variant reproduces."""
    o = refnodes
    rng = np.random.default_rng(33)
    seen = dict(far=0, invalid=0, low_s=0, ok=0)
    for trial in range(6):
        noise = [0.005, 0.03, 0.08, 0.15, 0.02, 0.05][trial]
        n_map = [30000, 30000, 20000, 20000, 4000, 8000][trial]
        xy = rng.uniform(-30, 30, size=(n_map, 2))
        planes = np.where(rng.random(n_map) < 0.5, -1.7 + noise * rng.normal(size=n_map), 0.0)
        wall = rng.random(n_map) < 0.25
        mp = np.zeros((n_map, 4), np.float32)
        mp[:, 0] = np.where(wall, 12.0 + noise * rng.normal(size=n_map), xy[:, 0]); mp[:, 1] = xy[:, 1]
        mp[:, 2] = np.where(wall, rng.uniform(-1.7, 4.0, n_map), planes + np.where(planes == 0.0, rng.uniform(-1.7, 6.0, n_map), 0.0))
        mp = o.voxel_grid(mp, 0.5)[0]
        n_q = 3000
        q = np.zeros((n_q, 4), np.float32)
        q[:, :2] = rng.uniform(-28, 28, size=(n_q, 2)) * (0.05 if trial == 5 else 1.0); q[:, 2] = np.where(rng.random(n_q) < 0.6, -1.7, rng.uniform(-1.7, 5.0, n_q)) + 0.3 * rng.normal(size=n_q) * (trial % 2)
        tf = np.array([0.01, -0.02, 0.03, 0.2, -0.1, 0.05], np.float32) * (trial - 2)
        R = o.RefMapOpt()
        R.set_scan_and_map(q, mp)
        R.set_transform(tf)
        coeff, flag = R.surf_optimization()
        R.close()
        so = o.surf_optimization(q, mp, tf)
        kidx, kd2 = o.kdtree_knn5(mp, so["sel"])
        assert np.array_equal(flag, so["flag"]), (trial, int(flag.sum()), int(so["flag"].sum()))
        assert np.array_equal(_bits(coeff[flag == 1]), _bits(so["coeff"][flag == 1])), trial
        far = kd2[:, 4] >= 1.0
        seen["far"] += int(far.sum()); seen["ok"] += int(flag.sum())
        seen["invalid"] += int(((~far) & (so["plane"][:, :3] == 0).all(1) & (flag == 0)).sum())
        seen["low_s"] += int(((~far) & (flag == 0)).sum())
    assert seen["far"] > 1000 and seen["ok"] > 3000 and seen["low_s"] > 500, seen


def test_image_projection_fuzzed_filters_and_rates(refnodes, synth):
    """random ParamServer settings through the reference's ImageProjection node: downsampleRate 1-4, point_filter_num 1-7, range windows that cut returns on both sides,
    N_SCAN below the sensor's ring count (the ring gate, :585), IMU rates of 100 / 200 / 500 Hz with jitter, scans whose last points lie beyond the last IMU row (findRotation
    never extrapolates, trap 3) — the published cloud equals the oracle's bit for bit, and the kept-index list equals the raw-index filter (trap 1)."""
    o = refnodes
    rng = np.random.default_rng(44)
    total = 0
    for trial in range(10):
        filt = dict(lidarMinRange=float(rng.choice([0.5, 1.0, 3.0, 6.0])), lidarMaxRange=float(rng.choice([25.0, 60.0, 1000.0])), N_SCAN=int(rng.choice([64, 48, 32])),
                    downsampleRate=int(rng.integers(1, 5)), point_filter_num=int(rng.integers(1, 8)))
        rate = float(rng.choice([100.0, 200.0, 500.0]))
        omega = tuple(rng.normal(scale=[0.05, 0.05, 0.9]))
        R = o.RefImageProjection(imuRate=rate, **filt)
        cur = 500.0 + float(rng.uniform(0, 0.01))
        raws = [synth.scan(synth.HDL64, np.array([0, 0, 0, 1.0 * k, 0, 0], np.float64), omega=omega, seed=1000 + 10 * trial + k) for k in range(3)]
        end = cur + float(raws[0]["time"][-1])
        t_hi = end + 0.3 if trial % 3 else end + 0.004                       # every third trial: the IMU stream stops right after the scan end (few rows past it)
        stamps, gyro = _imu_stream(rng, cur - 0.3, t_hi, rate, omega)
        for s_, g_ in zip(stamps, gyro):
            R.imu(s_, g_)
        for k in range(3):
            n = R.cloud(cur + 0.1 * k, raws[k])
        covered = stamps[0] <= cur and stamps[-1] >= end
        assert n == (1 if covered else 0), (trial, n, covered)
        if not covered:
            R.close(); continue
        info = R.last_info()
        od = o.imu_deskew_info(stamps, gyro, cur, end)
        assert info["imuAvailable"] == int(od["available"])
        cloud, kept = o.project_point_cloud(raws[0], filt, cur, od["imu_time"], od["imu_rot"], od["imu_pointer_cur"], bool(od["available"]))
        assert info["cloud"].shape == cloud.shape and np.array_equal(_bits(info["cloud"]), _bits(cloud)), (trial, filt)
        r0 = raws[0]
        rng_ok = np.sqrt((r0["x"].astype(np.float32) ** 2 + r0["y"].astype(np.float32) ** 2 + r0["z"].astype(np.float32) ** 2).astype(np.float32))
        want = np.nonzero((rng_ok >= filt["lidarMinRange"]) & (rng_ok <= filt["lidarMaxRange"]) & (r0["ring"] < filt["N_SCAN"]) & (r0["ring"] % filt["downsampleRate"] == 0)
                          & (np.arange(len(r0)) % filt["point_filter_num"] == 0))[0]
        assert abs(len(want) - len(kept)) <= 2 and np.all(kept % filt["point_filter_num"] == 0)      # (float range at the window edges may differ by an ulp from numpy's)
        total += len(kept)
        R.close()
    assert total > 20000


def test_lm_optimization_fuzzed_against_the_reference(refnodes):
    """LMOptimization (:1158-1293) on synthetic correspondences: well-conditioned, rank-deficient (all normals parallel → five small eigenvalues), nearly degenerate, fewer than 50
    rows, converged-size steps; iteration 0 (eigen-analysis, matP) followed by later iterations that reuse isDegenerate / matP — pose, verdict, isDegenerate and matP bits."""
    o = refnodes
    rng = np.random.default_rng(55)
    outcomes = dict(conv=0, deg=0, short=0)
    for trial in range(40):
        n = int(rng.choice([10, 49, 50, 51, 300, 2000, 6000]))
        kind = trial % 4
        ori = np.zeros((n, 4), np.float32); ori[:, :3] = rng.normal(scale=[20, 20, 3], size=(n, 3))
        nrm = rng.normal(size=(n, 3))
        if kind == 1:
            nrm = np.tile([0.0, 0.0, 1.0], (n, 1)) + 1e-4 * rng.normal(size=(n, 3))          # ground only
        elif kind == 2:
            nrm[:, 2] *= 1e-3                                                                # walls only: z unobservable
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        res = rng.normal(scale=[0.05, 1e-4, 0.02, 0.2][kind], size=n)
        s = rng.uniform(0.1, 1.0, size=n)
        coeff = np.concatenate([s[:, None] * nrm, (s * res)[:, None]], 1).astype(np.float32)
        tf = rng.normal(scale=[0.02, 0.02, 1.0, 20, 20, 1], size=6).astype(np.float32)
        R = o.RefMapOpt()
        st = np.zeros(37, np.float32)
        rtf = tf.copy()
        for it in range(3):
            conv, rtf = R.lm_optimization(it, ori, coeff, rtf)
            r = o.lm_optimization(it, ori, coeff, tf, st)
            tf, st = r["tf"], r["state"]
            assert conv == r["converged"] and np.array_equal(_bits(rtf), _bits(tf)), (trial, it, n, kind)
            deg, P = R.lm_state()
            assert deg == int(st[0]) and np.array_equal(_bits(P), _bits(st[1:])), (trial, it)
            outcomes["conv"] += int(conv); outcomes["short"] += int(n < 50)
        outcomes["deg"] += int(st[0])
        R.close()
    assert outcomes["deg"] >= 8 and outcomes["short"] >= 9 and outcomes["conv"] >= 3, outcomes


def test_keyframe_selection_adversarial_vs_reference(refnodes):
    """liorf_host_extract_nearby's grid-pruned nearest-1 (exact by construction: own-voxel bound, cell cover, shell bound, full-scan fallback) against the reference's
    extractNearby on data built to break a shortcut: key poses on an exact 1 m lattice (voxel centroids equidistant from several poses: ties), exact duplicates, a block
    driven round many times (hundreds of poses inside the search ball), far-away clusters, poses just inside / outside the radius, several radii and densities."""
    import liorf_b200
    o = refnodes
    vp, extract_nearby, _ = _host(liorf_b200.load_library())
    rng = np.random.default_rng(66)

    def check(poses, times, t_cur, radius, density):
        poses = np.ascontiguousarray(poses, np.float32)
        R = o.RefMapOpt(surroundingKeyframeSearchRadius=radius, surroundingKeyframeDensity=density)
        for k in range(len(poses)):
            R.add_keyframe(np.array([[0.5, 0.0, 0.0, float(k)]], np.float32), poses[k], times[k])
        R.extract_surrounding_keyframes(t_cur)
        got = R.get_cloud(2)[:, 3].astype(np.int64)
        R.close()
        want = extract_nearby(poses, times, t_cur, radius, density)
        assert np.array_equal(got, want), (len(poses), radius, density, got[:10], want[:10])
        return len(want)

    total = 0
    for trial in range(14):
        n = int(rng.choice([40, 150, 400, 900]))
        kind = trial % 7
        P = np.zeros((n, 6), np.float32)
        if kind == 0:                                    # exact lattice, visited in a random order
            g = np.stack(np.meshgrid(np.arange(30), np.arange(30), [0.0]), -1).reshape(-1, 3)[rng.permutation(900)[:n]]
            P[:, 3:6] = g
        elif kind == 1:                                  # duplicates of a few positions
            base = rng.uniform(-30, 30, size=(12, 3)) * [1, 1, 0.05]
            P[:, 3:6] = base[rng.integers(0, 12, n)]
        elif kind == 2:                                  # one block, many laps (1.3 m steps with jitter)
            s = np.arange(n) * 1.3
            per = 160.0; u = s % per
            P[:, 3] = np.where(u < 50, u, np.where(u < 80, 50, np.where(u < 130, 130 - u, 0))) + rng.normal(scale=0.05, size=n)
            P[:, 4] = np.where(u < 50, 0, np.where(u < 80, u - 50, np.where(u < 130, 30, 160 - u))) + rng.normal(scale=0.05, size=n)
        elif kind == 3:                                  # clusters, some far away
            c = rng.uniform(-300, 300, size=(6, 3)) * [1, 1, 0.01]; c[0] = 0
            P[:, 3:6] = c[rng.integers(0, 6, n)] + rng.normal(scale=8.0, size=(n, 3)) * [1, 1, 0.05]
        elif kind == 4:                                  # a ring of poses right at the radius around the newest one
            ang = rng.uniform(0, 2 * np.pi, n); rad = 50.0 + rng.choice([-1e-3, 0.0, 1e-3, -2.0, 2.0], n)
            P[:, 3] = rad * np.cos(ang); P[:, 4] = rad * np.sin(ang); P[-1, 3:6] = 0
        elif kind == 5:                                  # straight drive with exact 0.5 m spacing (pairs share 2 m voxels symmetrically)
            P[:, 3] = 0.5 * np.arange(n)
        else:                                            # random walk
            P[:, 3:6] = np.cumsum(rng.normal(scale=[1.0, 0.6, 0.02], size=(n, 3)), 0)
        times = 100.0 + 0.1 * np.arange(n) * rng.choice([1.0, 3.0])
        for radius, density in ((50.0, 2.0), (15.0, 1.0), (80.0, 5.0)):
            total += check(P, times, times[-1] + float(rng.choice([0.05, 5.0, 50.0])), radius, density)
    assert total > 3000

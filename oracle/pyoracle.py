"""ctypes binding of the CPU oracle (oracle/libliorf_oracle.so, oracle/_ref/libliorf_ref.so).

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module.  The product package (liorf_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
P4 = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("i", "<f4")])
PRAW = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("i", "<f4"), ("ring", "<u2"), ("pad", "<u2"), ("time", "<f4")])


def build(force=False):
    so = os.path.join(HERE, "libliorf_oracle.so")
    srcs = [os.path.join(HERE, f) for f in ("oracle_capi.cpp", "liorf_oracle.hpp", "ref_nanoflann.cpp", "ref_scancontext.cpp", "ref_mapopt.cpp", "ref_imageproj.cpp", "Makefile")]
    stale = force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs if os.path.exists(s))
    if stale:
        subprocess.run(["make", "-C", HERE, "-s"], check=True)
    return so


def _fp(a):
    return a.ctypes.data_as(C.c_void_p)


def _as_p4(a):
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 4)
    return a


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_sc_create.restype = C.c_void_p
        _lib.orc_get_transformation.argtypes = [C.c_float] * 6 + [C.c_void_p]
    return _lib


def ref():
    """The reference's vendored nanoflann compiled into oracle/_ref (None if it was never built)."""
    global _ref
    if _ref is None:
        build()
        p = os.path.join(HERE, "_ref", "libliorf_ref.so")
        if not os.path.exists(p):
            return None
        _ref = C.CDLL(p)
        _ref.ref_ringkey_tree_build_seconds.restype = C.c_double
    return _ref


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


def get_transformation(x, y, z, roll, pitch, yaw):
    t = np.zeros(12, np.float32)
    lib().orc_get_transformation(x, y, z, roll, pitch, yaw, _fp(t))
    return t.reshape(3, 4)


def transform_cloud(pts, pose6):
    pts = _as_p4(pts)
    out = np.empty_like(pts)
    pose6 = np.ascontiguousarray(pose6, np.float32)
    lib().orc_transform_cloud(_fp(pts), C.c_int(len(pts)), _fp(pose6), _fp(out))
    return out


def voxel_grid(pts, leaf, want_meta=False):
    pts = _as_p4(pts)
    n = len(pts)
    out = np.empty((max(n, 1), 4), np.float32)
    mem = np.empty(max(n, 1), np.int32)
    keys = np.empty(max(n, 1), np.int32)
    meta = np.zeros(6, np.int32)
    r = lib().orc_voxel_grid(_fp(pts), C.c_int(n), C.c_float(leaf), _fp(out), _fp(mem), _fp(keys), _fp(meta))
    if r < 0:
        res = (pts.copy(), np.arange(n, dtype=np.int32), None)
    else:
        res = (out[:r].copy(), mem[:n].copy(), keys[:r].copy())
    return res + (meta,) if want_meta else res


def knn5(map_pts, q):
    map_pts = _as_p4(map_pts); q = _as_p4(q)
    idx = np.empty((len(q), 5), np.int32); d2 = np.empty((len(q), 5), np.float32)
    lib().orc_knn5(_fp(map_pts), C.c_int(len(map_pts)), _fp(q), C.c_int(len(q)), _fp(idx), _fp(d2))
    return idx, d2


def surf_optimization(scan, map_pts, tf6):
    scan = _as_p4(scan); map_pts = _as_p4(map_pts); tf6 = np.ascontiguousarray(tf6, np.float32)
    n = len(scan)
    coeff = np.zeros((n, 4), np.float32); flag = np.zeros(n, np.uint8)
    idx = np.empty((n, 5), np.int32); d2 = np.empty((n, 5), np.float32); plane = np.zeros((n, 4), np.float32)
    sel = np.zeros((n, 4), np.float32)
    lib().orc_surf_optimization(_fp(scan), C.c_int(n), _fp(map_pts), C.c_int(len(map_pts)), _fp(tf6), _fp(coeff), _fp(flag),
                                _fp(idx), _fp(d2), _fp(plane), _fp(sel))
    return dict(coeff=coeff, flag=flag, idx=idx, d2=d2, plane=plane, sel=sel)


def combine(scan, coeff, flag):
    scan = _as_p4(scan); coeff = _as_p4(coeff); flag = np.ascontiguousarray(flag, np.uint8)
    ori = np.empty_like(scan); cs = np.empty_like(coeff)
    k = lib().orc_combine_optimization_coeffs(_fp(scan), _fp(coeff), _fp(flag), C.c_int(len(scan)), _fp(ori), _fp(cs))
    return ori[:k].copy(), cs[:k].copy()


def lm_optimization(it, ori, coeff, tf6, state37=None):
    ori = _as_p4(ori); coeff = _as_p4(coeff)
    tf = np.array(tf6, np.float32).copy()
    st = np.zeros(37, np.float32) if state37 is None else np.array(state37, np.float32).copy()
    tr = np.zeros(52, np.float32)
    c = lib().orc_lm_optimization(C.c_int(it), _fp(ori), _fp(coeff), C.c_int(len(ori)), _fp(tf), _fp(st), _fp(tr))
    return dict(converged=bool(c), tf=tf, state=st, AtA=tr[:36].reshape(6, 6).copy(), AtB=tr[36:42].copy(), X=tr[42:48].copy(),
                nsel=int(tr[48]), degenerate=bool(tr[50]), solved=bool(tr[51]))


def scan2map(scan, map_pts, tf6, max_iters=30, force_all=False, state37=None, use_ref_kdtree=False):
    scan = _as_p4(scan); map_pts = _as_p4(map_pts)
    tf = np.array(tf6, np.float32).copy()
    st = np.zeros(37, np.float32) if state37 is None else np.array(state37, np.float32).copy()
    trace = np.zeros((max_iters, 6), np.float32); nsel = np.zeros(max_iters, np.int32)
    if use_ref_kdtree:
        tm = np.zeros(4, np.float64)
        it = ref().ref_scan2map(_fp(scan), C.c_int(len(scan)), _fp(map_pts), C.c_int(len(map_pts)), _fp(tf), C.c_int(max_iters),
                                C.c_int(int(force_all)), _fp(st), _fp(trace), _fp(nsel), _fp(tm))
        return dict(iters=it, tf=tf, state=st, trace=trace[:it].copy(), nsel=nsel[:it].copy(), timings=tm)
    it = lib().orc_scan2map(_fp(scan), C.c_int(len(scan)), _fp(map_pts), C.c_int(len(map_pts)), _fp(tf), C.c_int(max_iters),
                            C.c_int(int(force_all)), _fp(st), _fp(trace), _fp(nsel))
    return dict(iters=it, tf=tf, state=st, trace=trace[:it].copy(), nsel=nsel[:it].copy())


def update_initial_guess(state31, no_keyframes, cloud_info11, use_imu_heading=True, imu_type=1):
    """state31 = tf[6] + lastImuTransformation[12] + lastImuPreTransformation[12] + available; cloud_info11 = (imuAvailable,
    odomAvailable, imuRollInit, imuPitchInit, imuYawInit, initialGuessX, Y, Z, Roll, Pitch, Yaw).  Returns the new state."""
    st = np.array(state31, np.float32).copy(); ci = np.ascontiguousarray(cloud_info11, np.float32)
    lib().orc_update_initial_guess(_fp(st), C.c_int(int(no_keyframes)), _fp(ci), C.c_int(int(use_imu_heading)), C.c_int(int(imu_type)))
    return st


def transform_update(tf6, imu_available, imu_type, imu_roll, imu_pitch, weight, rot_tol, z_tol):
    tf = np.array(tf6, np.float32).copy()
    lib().orc_transform_update(_fp(tf), C.c_int(int(imu_available)), C.c_int(int(imu_type)), C.c_float(imu_roll), C.c_float(imu_pitch), C.c_float(weight),
                               C.c_float(rot_tol), C.c_float(z_tol))
    return tf


def project_point_cloud(raw, params, time_scan_cur, imu_time, imu_rot, imu_pointer_cur, deskew_enabled=True):
    """raw: structured PRAW array; imu_rot: (rows,3) float64; params: dict(lidarMinRange,lidarMaxRange,N_SCAN,downsampleRate,point_filter_num)."""
    raw = np.ascontiguousarray(raw, dtype=PRAW)
    n = len(raw)
    out = np.empty((max(n, 1), 4), np.float32); kept = np.empty(max(n, 1), np.int32)
    it = np.ascontiguousarray(imu_time, np.float64)
    rot = np.ascontiguousarray(imu_rot, np.float64)
    rx, ry, rz = (np.ascontiguousarray(rot[:, k]) for k in range(3))
    r = lib().orc_project_point_cloud(_fp(raw), C.c_int(n), C.c_float(params["lidarMinRange"]), C.c_float(params["lidarMaxRange"]),
                                      C.c_int(params["N_SCAN"]), C.c_int(params["downsampleRate"]), C.c_int(params["point_filter_num"]),
                                      C.c_double(time_scan_cur), _fp(it), _fp(rx), _fp(ry), _fp(rz), C.c_int(imu_pointer_cur),
                                      C.c_int(int(deskew_enabled)), _fp(out), _fp(kept))
    return out[:r].copy(), kept[:r].copy()


def imu_deskew_info(stamp, gyro_xyz, time_scan_cur, time_scan_end):
    """oracle restatement of imuDeskewInfo (src/imageProjection.cpp:350-409) over a std::deque; at most 2000 rows"""
    stamp = np.ascontiguousarray(stamp, np.float64); g = np.ascontiguousarray(gyro_xyz, np.float64).reshape(-1, 3)
    t = np.zeros(2000); rx = np.zeros(2000); ry = np.zeros(2000); rz = np.zeros(2000); ptr = C.c_int(-1); left = C.c_int(0)
    ok = lib().orc_imu_deskew_info(_fp(stamp), _fp(g), C.c_int(len(stamp)), C.c_double(time_scan_cur), C.c_double(time_scan_end), _fp(t), _fp(rx), _fp(ry), _fp(rz),
                                   C.byref(ptr), C.byref(left))
    rows = max(ptr.value + 1, 0)
    return dict(available=bool(ok), imu_time=t[:rows].copy(), imu_rot=np.stack([rx[:rows], ry[:rows], rz[:rows]], 1), imu_pointer_cur=ptr.value, queue_left=left.value)


def sc_make(pts):
    pts = _as_p4(pts)
    desc = np.zeros(1200, np.float64); rk = np.zeros(20, np.float32); sk = np.zeros(60, np.float64)
    lib().orc_sc_make(_fp(pts), C.c_int(len(pts)), _fp(desc), _fp(rk), _fp(sk))
    return desc.reshape(20, 60), rk, sk


def sc_keys_from_desc(desc):
    desc = np.ascontiguousarray(desc, np.float64).reshape(-1)
    rk = np.zeros(20, np.float32); sk = np.zeros(60, np.float64)
    lib().orc_sc_keys_from_desc(_fp(desc), _fp(rk), _fp(sk))
    return rk, sk


def sc_distance(sc1, sc2):
    sc1 = np.ascontiguousarray(sc1, np.float64).reshape(-1); sc2 = np.ascontiguousarray(sc2, np.float64).reshape(-1)
    d = C.c_double(); s = C.c_int()
    lib().orc_sc_distance(_fp(sc1), _fp(sc2), C.byref(d), C.byref(s))
    return d.value, s.value


def ringkey_top3(keys, q, use_ref=False):
    keys = np.ascontiguousarray(keys, np.float32).reshape(-1, 20); q = np.ascontiguousarray(q, np.float32).reshape(-1, 20)
    idx = np.zeros((len(q), 3), np.int32); d = np.zeros((len(q), 3), np.float32)
    f = ref().ref_ringkey_knn3 if use_ref else lib().orc_ringkey_top3
    f(_fp(keys), C.c_int(len(keys)), _fp(q), C.c_int(len(q)), _fp(idx), _fp(d))
    return idx, d


def kdtree_knn5(map_pts, q):
    map_pts = _as_p4(map_pts); q = _as_p4(q)
    idx = np.empty((len(q), 5), np.int32); d2 = np.empty((len(q), 5), np.float32)
    ref().ref_kdtree_knn5(_fp(map_pts), C.c_int(len(map_pts)), _fp(q), C.c_int(len(q)), _fp(idx), _fp(d2))
    return idx, d2


def sc_query_batch(keys, descs, qkeys, qdescs):
    keys = np.ascontiguousarray(keys, np.float32).reshape(-1, 20); descs = np.ascontiguousarray(descs, np.float64).reshape(-1, 1200)
    qkeys = np.ascontiguousarray(qkeys, np.float32).reshape(-1, 20); qdescs = np.ascontiguousarray(qdescs, np.float64).reshape(-1, 1200)
    Q = len(qkeys)
    loop = np.zeros(Q, np.int32); shift = np.zeros(Q, np.int32); dist = np.zeros(Q, np.float64); cand = np.zeros((Q, 3), np.int32)
    lib().orc_sc_query_batch(_fp(keys), _fp(descs), C.c_int(len(keys)), _fp(qkeys), _fp(qdescs), C.c_int(Q), _fp(loop), _fp(shift), _fp(dist), _fp(cand))
    return loop, shift, dist, cand


def ref_sc_query_batch(keys, descs, qkeys, qdescs):
    """the same batch with the reference's vendored nanoflann kd-tree as stage 1 (oracle/_ref); returns (loop, shift, dist, cand, (tree_s, query_s))"""
    keys = np.ascontiguousarray(keys, np.float32).reshape(-1, 20); descs = np.ascontiguousarray(descs, np.float64).reshape(-1, 1200)
    qkeys = np.ascontiguousarray(qkeys, np.float32).reshape(-1, 20); qdescs = np.ascontiguousarray(qdescs, np.float64).reshape(-1, 1200)
    Q = len(qkeys)
    loop = np.zeros(Q, np.int32); shift = np.zeros(Q, np.int32); dist = np.zeros(Q, np.float64); cand = np.zeros((Q, 3), np.int32); tm = np.zeros(2, np.float64)
    ref().ref_sc_query_batch(_fp(keys), _fp(descs), C.c_int(len(keys)), _fp(qkeys), _fp(qdescs), C.c_int(Q), _fp(loop), _fp(shift), _fp(dist), _fp(cand), _fp(tm))
    return loop, shift, dist, cand, (float(tm[0]), float(tm[1]))


def sc_keys_batch(descs):
    """ring keys (fp32) of many descriptors"""
    descs = np.ascontiguousarray(descs, np.float64).reshape(-1, 1200)
    return np.stack([sc_keys_from_desc(d)[0] for d in descs]) if len(descs) else np.zeros((0, 20), np.float32)


class SCManager:
    def __init__(self):
        self.h = C.c_void_p(lib().orc_sc_create())

    def __del__(self):
        try:
            lib().orc_sc_destroy(self.h)
        except Exception:
            pass

    def make_and_save(self, pts):
        pts = _as_p4(pts)
        lib().orc_sc_make_and_save(self.h, _fp(pts), C.c_int(len(pts)))

    def save_descriptor(self, desc):
        desc = np.ascontiguousarray(desc, np.float64).reshape(-1)
        lib().orc_sc_save_descriptor(self.h, _fp(desc))

    def size(self):
        return lib().orc_sc_size(self.h)

    def get(self, i):
        d = np.zeros(1200, np.float64); k = np.zeros(20, np.float32)
        lib().orc_sc_get(self.h, C.c_int(i), _fp(d), _fp(k))
        return d.reshape(20, 60), k

    def detect(self):
        lid = C.c_int(); yaw = C.c_float(); md = C.c_double(); cand = np.zeros(3, np.int32)
        lib().orc_sc_detect(self.h, C.byref(lid), C.byref(yaw), C.byref(md), _fp(cand))
        return lid.value, yaw.value, md.value, cand


_refsc = None


def refsc():
    """The reference's OWN include/Scancontext.cpp compiled unchanged against oracle/shim (oracle/_ref/libliorf_ref_sc.so); None if never built."""
    global _refsc
    if _refsc is None:
        build()
        p = os.path.join(HERE, "_ref", "libliorf_ref_sc.so")
        if not os.path.exists(p):
            return None
        _refsc = C.CDLL(p)
        _refsc.refsc_create.restype = C.c_void_p
        _refsc.refsc_xy2theta.restype = C.c_float
        _refsc.refsc_xy2theta.argtypes = [C.c_float, C.c_float]
    return _refsc


class RefSCManager:
    """SCManager of the reference itself (include/Scancontext.h:61-118) — the pin of the oracle's restatement"""

    def __init__(self):
        self.h = C.c_void_p(refsc().refsc_create())

    def __del__(self):
        try:
            refsc().refsc_destroy(self.h)
        except Exception:
            pass

    def make(self, pts):
        pts = _as_p4(pts)
        desc = np.zeros(1200, np.float64); rk = np.zeros(20, np.float32); sk = np.zeros(60, np.float64)
        refsc().refsc_make(self.h, _fp(pts), C.c_int(len(pts)), _fp(desc), _fp(rk), _fp(sk))
        return desc.reshape(20, 60), rk, sk

    def make_and_save(self, pts):
        pts = _as_p4(pts)
        refsc().refsc_make_and_save(self.h, _fp(pts), C.c_int(len(pts)))

    def save_descriptor(self, desc):
        desc = np.ascontiguousarray(desc, np.float64).reshape(-1)
        refsc().refsc_save_descriptor(self.h, _fp(desc))

    def size(self):
        return refsc().refsc_size(self.h)

    def get(self, i):
        d = np.zeros(1200, np.float64); k = np.zeros(20, np.float32)
        refsc().refsc_get(self.h, C.c_int(i), _fp(d), _fp(k))
        return d.reshape(20, 60), k

    def detect(self):
        lid = C.c_int(); yaw = C.c_float()
        refsc().refsc_detect(self.h, C.byref(lid), C.byref(yaw))
        return lid.value, yaw.value

    def query_batch(self, qdescs):
        """every query through the reference's own detectLoopClosureID, unchanged (see refsc_query_batch_own) → (loop_id, shift)"""
        qd = np.ascontiguousarray(qdescs, np.float64).reshape(-1, 1200)
        loop = np.zeros(len(qd), np.int32); sh = np.zeros(len(qd), np.int32)
        refsc().refsc_query_batch_own(self.h, _fp(qd), C.c_int(len(qd)), _fp(loop), _fp(sh))
        return loop, sh

    def save_descriptors(self, descs):
        d = np.ascontiguousarray(descs, np.float64).reshape(-1, 1200)
        f = refsc().refsc_save_descriptor
        for k in range(len(d)):
            f(self.h, C.c_void_p(d[k].ctypes.data))

    def distance(self, sc1, sc2):
        sc1 = np.ascontiguousarray(sc1, np.float64).reshape(-1); sc2 = np.ascontiguousarray(sc2, np.float64).reshape(-1)
        d = C.c_double(); s = C.c_int()
        refsc().refsc_distance(self.h, _fp(sc1), _fp(sc2), C.byref(d), C.byref(s))
        return d.value, s.value


class RefMapOpt:
    """The reference's OWN mapOptimization node (src/mapOptmization.cpp compiled unchanged against oracle/shim_ros into
    oracle/_ref/libliorf_ref_mapopt.so) — the pin of the oracle's a4-a9 / f1 / f2 restatements.  The node keeps function-static state
    (timeLastProcessing :254, lastImuTransformation :904 ...), so every instance loads its own private copy of the library.
    params: ParamServer names without the "liorf/" prefix (include/utility.h:153-237)."""

    @staticmethod
    def available():
        build()
        return os.path.exists(os.path.join(HERE, "_ref", "libliorf_ref_mapopt.so"))

    def __init__(self, openmp=False, **params):
        """openmp=True loads the build with the reference's own flags (-O3, OpenMP: bench.py's CPU arm; surfOptimization then writes its std::vector<bool>
        flags from several threads as the reference does, :1078,1137); the default build has no OpenMP (tests: deterministic)."""
        import shutil
        import tempfile
        build()
        src = os.path.join(HERE, "_ref", "libliorf_ref_mapopt_omp.so" if openmp else "libliorf_ref_mapopt.so")
        fd, self._path = tempfile.mkstemp(suffix=".so", prefix="liorf_ref_mapopt_")
        os.close(fd)
        shutil.copyfile(src, self._path)
        self.l = C.CDLL(self._path)
        self.l.refmo_create.restype = C.c_void_p
        self.l.refmo_param_num.argtypes = [C.c_char_p, C.c_double]
        self.l.refmo_param_str.argtypes = [C.c_char_p, C.c_char_p]
        defaults = dict(sensor="velodyne", N_SCAN=64, Horizon_SCAN=1800, mappingProcessInterval=0.0, numberOfCores=1, mappingSurfLeafSize=0.4,
                        surroundingKeyframeMapLeafSize=0.5, surroundingKeyframeDensity=2.0, surroundingKeyframeSearchRadius=50.0,
                        surroundingkeyframeAddingDistThreshold=1.0, surroundingkeyframeAddingAngleThreshold=0.2, z_tollerance=1000.0,
                        rotation_tollerance=1000.0, imuType=0, imuRPYWeight=0.01)      # config/kitti.yaml
        defaults.update(params)
        for k, v in defaults.items():
            if isinstance(v, str):
                self.l.refmo_param_str(("liorf/" + k).encode(), v.encode())
            else:
                self.l.refmo_param_num(("liorf/" + k).encode(), float(v))
        self.h = C.c_void_p(self.l.refmo_create())

    def close(self):
        if self.h:
            self.l.refmo_destroy(self.h); self.h = None
            try:
                os.unlink(self._path)
            except OSError:
                pass

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def cloud_info(self, stamp, cloud, imu_available=0, odom_available=0, rpy_init=None, guess=None):
        """one cloud_info message through laserCloudInfoHandler (:236-275)"""
        cloud = _as_p4(cloud)
        r = None if rpy_init is None else np.ascontiguousarray(rpy_init, np.float32)
        g = None if guess is None else np.ascontiguousarray(guess, np.float32)
        self.l.refmo_cloud_info(self.h, C.c_double(stamp), _fp(cloud), C.c_int(len(cloud)), C.c_int(int(imu_available)), C.c_int(int(odom_available)),
                                None if r is None else _fp(r), None if g is None else _fp(g))

    def state(self):
        tf = np.zeros(6, np.float32); cnt = np.zeros(5, np.int32)
        self.l.refmo_state(self.h, _fp(tf), _fp(cnt))
        return dict(tf=tf, keyframes=int(cnt[0]), n_ds=int(cnt[1]), m_ds=int(cnt[2]), degenerate=int(cnt[3]), sc_entries=int(cnt[4]))

    def set_transform(self, tf6):
        tf = np.ascontiguousarray(tf6, np.float32)
        self.l.refmo_set_transform(self.h, _fp(tf))

    def get_cloud(self, which):
        n = self.l.refmo_get_cloud(self.h, C.c_int(which), None, C.c_int(0))
        if n < 0:
            return None
        out = np.zeros((max(n, 1), 4), np.float32)
        self.l.refmo_get_cloud(self.h, C.c_int(which), _fp(out), C.c_int(n))
        return out[:n].copy()

    def keypose(self, k):
        p = np.zeros(6, np.float32); t = C.c_double()
        if self.l.refmo_get_keypose(self.h, C.c_int(k), _fp(p), C.byref(t)) < 0:
            return None
        return p, t.value

    def set_scan_and_map(self, scan_ds, map_ds):
        a = _as_p4(scan_ds); b = _as_p4(map_ds)
        self.l.refmo_set_scan_and_map(self.h, _fp(a), C.c_int(len(a)), _fp(b), C.c_int(len(b)))
        self._n = len(a)

    def iteration(self, it):
        """{laserCloudOri/coeffSel clear, surfOptimization, combineOptimizationCoeffs, LMOptimization(it)} (:1304-1314) → (converged, tf, n_sel)"""
        tf = np.zeros(6, np.float32); ns = C.c_int()
        c = self.l.refmo_iteration(self.h, C.c_int(it), _fp(tf), C.byref(ns))
        return bool(c), tf, ns.value

    def surf_optimization(self):
        coeff = np.zeros((self._n, 4), np.float32); flag = np.zeros(self._n, np.uint8)
        self.l.refmo_surf_optimization(self.h, _fp(coeff), _fp(flag))
        return coeff, flag

    def scan2map(self, tf6, imu_available=0, rpy_init=None):
        tf = np.array(tf6, np.float32).copy()
        r = None if rpy_init is None else np.ascontiguousarray(rpy_init, np.float32)
        self.l.refmo_scan2map(self.h, _fp(tf), C.c_int(int(imu_available)), None if r is None else _fp(r))
        return tf

    def lm_state(self):
        P = np.zeros(36, np.float32)
        d = self.l.refmo_get_lm_state(self.h, _fp(P))
        return d, P

    def sc_detect(self):
        lid = C.c_int(); yaw = C.c_float()
        self.l.refmo_sc_detect(self.h, C.byref(lid), C.byref(yaw))
        return lid.value, yaw.value

    def set_map(self, map_ds):
        b = _as_p4(map_ds)
        self.l.refmo_set_map(self.h, _fp(b), C.c_int(len(b)))

    def bench_step(self, scan, tf6, iters=30, force_all=True):
        """one headline step on the reference's own member functions: downsampleCurrentScan + kd-tree build + `iters` passes of scan2MapOptimization's
        loop body.  Returns (final pose, iterations run, {downsample, kdtree_build, surf_optimization, lm_optimization} in ms)."""
        a = _as_p4(scan); tf = np.ascontiguousarray(tf6, np.float32); out = np.zeros(6, np.float32); tm = np.zeros(4, np.float64)
        it = self.l.refmo_bench_step(self.h, _fp(a), C.c_int(len(a)), _fp(tf), C.c_int(iters), C.c_int(int(force_all)), _fp(out), _fp(tm))
        return out, int(it), dict(downsample=float(tm[0]), kdtree_build=float(tm[1]), surf_optimization=float(tm[2]), lm_optimization=float(tm[3]))

    def add_keyframe(self, cloud, pose6, t):
        """harness set-up: stores a keyframe the way saveKeyFramesAndFactor does (:1548-1580)"""
        cloud = _as_p4(cloud); p = np.ascontiguousarray(pose6, np.float32)
        self.l.refmo_add_keyframe(self.h, _fp(cloud), C.c_int(len(cloud)), _fp(p), C.c_double(t))

    def _cloud_call(self, fn, *args):
        n = fn(self.h, *args, None, C.c_int(0))
        if n < 0:
            return None
        out = np.zeros((max(n, 1), 4), np.float32)
        fn(self.h, *args, _fp(out), C.c_int(n))
        return out[:n].copy()

    def lm_optimization(self, it, ori, coeff, tf6):
        """LMOptimization(it) (:1158-1293) on given (laserCloudOri, coeffSel) → (converged, tf)"""
        a = _as_p4(ori); b = _as_p4(coeff); tf = np.array(tf6, np.float32).copy()
        c = self.l.refmo_lm_optimization(self.h, C.c_int(it), _fp(a), _fp(b), C.c_int(len(a)), _fp(tf))
        return bool(c), tf

    def update_initial_guess(self, ci11, have_keyframes, tf6):
        """updateInitialGuess() (:899-958) on cloud_info = (imuAvailable, odomAvailable, imuRollInit, imuPitchInit, imuYawInit, initialGuessX, Y, Z, Roll, Pitch, Yaw)"""
        ci = np.ascontiguousarray(ci11, np.float32); tf = np.array(tf6, np.float32).copy()
        self.l.refmo_update_initial_guess(self.h, _fp(ci), C.c_int(int(have_keyframes)), _fp(tf))
        return tf

    def transform_update(self, ci11, tf6):
        ci = np.ascontiguousarray(ci11, np.float32); tf = np.array(tf6, np.float32).copy()
        self.l.refmo_transform_update(self.h, _fp(ci), _fp(tf))
        return tf

    def save_frame(self, last_pose6, tf6):
        last = None if last_pose6 is None else np.ascontiguousarray(last_pose6, np.float32); tf = np.ascontiguousarray(tf6, np.float32)
        return bool(self.l.refmo_save_frame(self.h, None if last is None else _fp(last), _fp(tf)))

    def extract_surrounding_keyframes(self, t):
        """extractSurroundingKeyFrames() at laser time t (:1046-1059); get_cloud(2) / get_cloud(1) then hold the raw / filtered local map"""
        self.l.refmo_extract_surrounding_keyframes(self.h, C.c_double(t))

    def loop_find_near_keyframes(self, key, search_num, loop_index):
        return self._cloud_call(self.l.refmo_loop_find_near_keyframes, C.c_int(key), C.c_int(search_num), C.c_int(loop_index))

    def publish_global_map(self):
        return self._cloud_call(self.l.refmo_publish_global_map)


class RefImageProjection:
    """The reference's OWN ImageProjection node (src/imageProjection.cpp compiled unchanged against oracle/shim_ros into
    oracle/_ref/libliorf_ref_imageproj.so), driven through its public handlers only: imu() = imuHandler, cloud() = cloudHandler; the published
    liorf/cloud_info is read back with last_info().  Every instance loads a private copy of the library (function-static state :291, :414)."""

    @staticmethod
    def available():
        build()
        return os.path.exists(os.path.join(HERE, "_ref", "libliorf_ref_imageproj.so"))

    def __init__(self, **params):
        import shutil
        import tempfile
        build()
        fd, self._path = tempfile.mkstemp(suffix=".so", prefix="liorf_ref_imageproj_")
        os.close(fd)
        shutil.copyfile(os.path.join(HERE, "_ref", "libliorf_ref_imageproj.so"), self._path)
        self.l = C.CDLL(self._path)
        self.l.refip_create.restype = C.c_void_p
        self.l.refip_cloud.restype = C.c_long
        self.l.refip_param_num.argtypes = [C.c_char_p, C.c_double]
        self.l.refip_param_str.argtypes = [C.c_char_p, C.c_char_p]
        defaults = dict(sensor="velodyne", N_SCAN=64, Horizon_SCAN=1800, downsampleRate=2, point_filter_num=5, lidarMinRange=1.0, lidarMaxRange=1000.0,
                        imuType=0, imuRate=100.0)                     # config/kitti.yaml
        defaults.update(params)
        for k, v in defaults.items():
            if isinstance(v, str):
                self.l.refip_param_str(("liorf/" + k).encode(), v.encode())
            else:
                self.l.refip_param_num(("liorf/" + k).encode(), float(v))
        self.h = C.c_void_p(self.l.refip_create())

    def close(self):
        if self.h:
            self.l.refip_destroy(self.h); self.h = None
            try:
                os.unlink(self._path)
            except OSError:
                pass

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def imu(self, stamp, gyro_xyz, quat_xyzw=None):
        g = np.ascontiguousarray(gyro_xyz, np.float64)
        q = None if quat_xyzw is None else np.ascontiguousarray(quat_xyzw, np.float64)
        self.l.refip_imu(self.h, C.c_double(stamp), _fp(g), None if q is None else _fp(q))

    def cloud(self, stamp, raw, has_time=True):
        """pushes one scan; returns the number of cloud_info messages published so far (the node answers two scans late, :209-215)"""
        raw = np.ascontiguousarray(raw, dtype=PRAW)
        return int(self.l.refip_cloud(self.h, C.c_double(stamp), _fp(raw), C.c_int(len(raw)), C.c_int(int(has_time))))

    def last_info(self):
        st = C.c_double(); av = np.zeros(2, np.int32); rpy = np.zeros(3, np.float32)
        n = self.l.refip_last_info(C.byref(st), _fp(av), _fp(rpy), None, C.c_int(0))
        if n < 0:
            return None
        out = np.zeros((max(n, 1), 4), np.float32)
        self.l.refip_last_info(C.byref(st), _fp(av), _fp(rpy), _fp(out), C.c_int(n))
        return dict(stamp=st.value, imuAvailable=int(av[0]), odomAvailable=int(av[1]), rpy_init=rpy, cloud=out[:n].copy())


def cv_qr_solve6(A, b):
    A = np.ascontiguousarray(A, np.float32).reshape(36); b = np.ascontiguousarray(b, np.float32).reshape(6)
    x = np.zeros(6, np.float32)
    ok = lib().orc_cv_qr_solve6(_fp(A), _fp(b), _fp(x))
    return x, bool(ok)


def cv_jacobi6(A):
    A = np.ascontiguousarray(A, np.float32).reshape(36)
    W = np.zeros(6, np.float32); V = np.zeros(36, np.float32)
    lib().orc_cv_jacobi6(_fp(A), _fp(W), _fp(V))
    return W, V.reshape(6, 6)


def cv_lu_invert6(A):
    A = np.ascontiguousarray(A, np.float32).reshape(36)
    inv = np.zeros(36, np.float32)
    ok = lib().orc_cv_lu_invert6(_fp(A), _fp(inv))
    return inv.reshape(6, 6), bool(ok)


def cv_gemm6(A, B):
    A = np.ascontiguousarray(A, np.float32).reshape(36); B = np.ascontiguousarray(B, np.float32).reshape(36)
    Cc = np.zeros(36, np.float32)
    lib().orc_cv_gemm6(_fp(A), _fp(B), _fp(Cc))
    return Cc.reshape(6, 6)


def colpiv_qr_solve_5x3(A, b):
    A = np.ascontiguousarray(A, np.float32).reshape(15); b = np.ascontiguousarray(b, np.float32).reshape(5)
    x = np.zeros(3, np.float32)
    lib().orc_colpiv_qr_solve_5x3(_fp(A), _fp(b), _fp(x))
    return x

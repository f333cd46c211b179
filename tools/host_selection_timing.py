"""Host cost of the keyframe selection (liorf_host_extract_nearby = extractNearby + extractCloud's gate, src/mapOptmization.cpp:975-1018) along the config-2 drive:
liorf_process_frame runs it once per frame (twice on keyframe frames with look-ahead), on the frame's critical path, so its cost must not grow with the number of key
poses faster than the O(n) radius scan.  Host-only: no GPU needed.   usage: python tools/host_selection_timing.py > profiles/r02_host_extract_nearby.txt"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_common as bc      # noqa: E402
import liorf_b200              # noqa: E402


def main():
    lib = liorf_b200.load_library()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    seq = bc.Sequence(2762, 0, loop=True)
    P = seq.poses
    kf = [0]
    for i in range(1, len(P)):                                       # the saveFrame gate (1 m / 0.2 rad) along the true trajectory
        if np.linalg.norm(P[i][3:] - P[kf[-1]][3:]) >= 1.0 or np.abs(P[i][:3] - P[kf[-1]][:3]).max() >= 0.2:
            kf.append(i)
    poses = np.ascontiguousarray(np.array([P[i] for i in kf], np.float32)); times = np.ascontiguousarray(100.0 + 0.1 * np.array(kf, np.float64))
    ids = np.zeros(8192, np.int32); cnt = C.c_int(0)
    print("config 2 drive: %d key poses; liorf_host_extract_nearby(radius 50 m, density 2 m), mean of 20 calls at each size (ctypes call included)" % len(kf))
    tot = 0.0; calls = 0
    for n in list(range(100, len(kf), 100)) + [len(kf)]:
        a = time.perf_counter()
        for _ in range(20):
            lib.liorf_host_extract_nearby(vp(poses), vp(times), n, C.c_double(times[n - 1] + 0.05), C.c_float(50.0), C.c_float(2.0), vp(ids), 8192, C.byref(cnt))
        us = (time.perf_counter() - a) / 20 * 1e6
        tot += us; calls += 1
        if n % 200 == 0 or n == len(kf):
            print("  %5d key poses: %6.1f us per call, %3d keyframes selected" % (n, us, cnt.value))
    print("mean over the sizes: %.1f us per call" % (tot / calls))
    return 0


if __name__ == "__main__":
    sys.exit(main())

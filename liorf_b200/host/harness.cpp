// harness.cpp — standalone C++ driver of the hot path through the host mirror (liorf_host.hpp) and the C ABI.
// Reads a binary case file written by tests (keyframes + scan + initial pose), runs
//   setLaserCloudSurfLast → downsampleCurrentScan → extractSurroundingKeyFrames → scan2MapOptimization → saveFrame
// and prints the pose; `tests/test_gpu_host_harness.py` compares it with the Python binding on the same bytes.
// File: int32 n_keyframes; per keyframe: int32 n, float pose6[6], double time, n x float4; int32 n_scan; n_scan x float4;
//       float init6[6]; double time_cur.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "liorf_host.hpp"

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: %s case.bin\n", argv[0]); return 2; }
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) { std::perror("open"); return 2; }
    auto rd = [&](void* p, size_t n) { if (std::fread(p, 1, n, f) != n) { std::fprintf(stderr, "short read\n"); std::exit(2); } };
    liorf_b200::ParamServer ps;
    liorf_b200::Context ctx(ps);
    liorf_b200::mapOptimization mo(ctx, ps);
    int nkf = 0; rd(&nkf, 4);
    for (int k = 0; k < nkf; ++k) {
        int n = 0; float pose[6]; double t; rd(&n, 4); rd(pose, 24); rd(&t, 8);
        std::vector<liorf_point> cloud(n); rd(cloud.data(), (size_t)n * 16);
        if (liorf_add_keyframe_cloud(ctx.get(), cloud.data(), n, pose, t) < 0) { std::fprintf(stderr, "add_keyframe failed\n"); return 1; }
    }
    int ns = 0; rd(&ns, 4);
    std::vector<liorf_point> scan(ns); rd(scan.data(), (size_t)ns * 16);
    rd(mo.transformTobeMapped, 24); rd(&mo.timeLaserInfoCur, 8);
    std::fclose(f);
    mo.setLaserCloudSurfLast(scan);
    mo.extractSurroundingKeyFrames();
    mo.downsampleCurrentScan();
    mo.scan2MapOptimization();
    const bool kf = mo.saveFrame();
    std::printf("pose %.9g %.9g %.9g %.9g %.9g %.9g iters %d converged %d degenerate %d save_frame %d\n", mo.transformTobeMapped[0], mo.transformTobeMapped[1],
                mo.transformTobeMapped[2], mo.transformTobeMapped[3], mo.transformTobeMapped[4], mo.transformTobeMapped[5], mo.lastTrace.iters,
                mo.lastTrace.converged, mo.lastTrace.degenerate, kf ? 1 : 0);
    return 0;
}

// ref_mapopt.cpp — extern "C" surface over the REFERENCE's own mapOptimization node: /root/reference/src/mapOptmization.cpp is compiled UNCHANGED,
// from where it lies (textually included below — the class has no header; nothing is copied into the repo), together with include/Scancontext.cpp and
// lib/common_lib.cpp, against the header stand-ins of oracle/shim_ros/ (ROS, PCL, Eigen, OpenCV, tf, GTSAM are not in this image).
// Output: oracle/_ref/libliorf_ref_mapopt.so (git-ignored, travels to the GPU box).  Recipe: oracle/Makefile.
//
// TEST INFRASTRUCTURE.  What it pins: the control flow, indexing, thresholds, float / double mixes, member state and call order of
// laserCloudInfoHandler -> updateInitialGuess / extractSurroundingKeyFrames (extractNearby, extractCloud) / downsampleCurrentScan / scan2MapOptimization
// (surfOptimization, combineOptimizationCoeffs, LMOptimization, transformUpdate) / saveKeyFramesAndFactor (saveFrame, keyframe store, ScanContext) are the
// reference's own code (src/mapOptmization.cpp:236-275, 899-1384, 1503-1609).  What it does NOT pin: the arithmetic inside the third-party calls, which the
// stand-ins forward to oracle/liorf_oracle.hpp (VoxelGrid, ColPivHouseholderQR, cv::solve / eigen / inv / products, getTransformation, Affine3f, tf), the
// kd-tree (the reference's vendored nanoflann instead of FLANN) and iSAM2 (stand-in: the estimate of a new pose is its initial value).  Compiled without
// OpenMP: the reference's surfOptimization writes a std::vector<bool> from several threads (src/mapOptmization.cpp:1078,1137), a race the test must not inherit.
#include <cmath>
#include <cstring>
#include <iostream>
#include <sstream>
#include <chrono>
#define main liorf_ref_mapopt_main_unused
#include "mapOptmization.cpp"
#undef main

namespace {
struct Quiet { std::streambuf* old; std::ostringstream sink; Quiet() : old(std::cout.rdbuf(sink.rdbuf())) {} ~Quiet() { std::cout.rdbuf(old); } };
void fill_cloud(const float* xyzi, int n, pcl::PointCloud<PointType>& c) {
    c.points.resize((size_t)n); c.width = (uint32_t)n;
    for (int i = 0; i < n; ++i) { PointType p; p.x = xyzi[4 * i]; p.y = xyzi[4 * i + 1]; p.z = xyzi[4 * i + 2]; p.intensity = xyzi[4 * i + 3]; c.points[i] = p; }
}
int copy_cloud(const pcl::PointCloud<PointType>& c, float* out, int cap) {
    const int n = (int)c.points.size();
    for (int i = 0; i < n && i < cap; ++i) { out[4 * i] = c.points[i].x; out[4 * i + 1] = c.points[i].y; out[4 * i + 2] = c.points[i].z; out[4 * i + 3] = c.points[i].intensity; }
    return n;
}
}  // namespace

extern "C" {

// parameters of ParamServer (include/utility.h:153-237), set BEFORE refmo_create; unset names take the reference's own defaults
void refmo_param_num(const char* name, double v) { ros::shim::params().num[name] = v; }
void refmo_param_str(const char* name, const char* v) { ros::shim::params().str[name] = v; }

void* refmo_create() {
    Quiet q;                                                       // common_lib prints a banner
    if (!ros::shim::params().str.count("liorf/sensor")) ros::shim::params().str["liorf/sensor"] = "velodyne";
    if (!common_lib_) common_lib_ = std::make_shared<CommonLib::common_lib>("mapping");                 // main() :1790
    return new mapOptimization();
}
void refmo_destroy(void* h) { delete (mapOptimization*)h; }

// one cloud_info message through laserCloudInfoHandler (:236-275).  cloud = the deskewed scan (x, y, z, intensity), rpy_init = cloudInfo.imu{Roll,Pitch,Yaw}Init,
// guess = cloudInfo.initialGuess{X,Y,Z,Roll,Pitch,Yaw}
void refmo_cloud_info(void* h, double stamp, const float* xyzi, int n, int imu_available, int odom_available, const float* rpy_init, const float* guess) {
    mapOptimization* mo = (mapOptimization*)h;
    auto msg = std::make_shared<liorf::cloud_info>();
    msg->header.stamp.fromSec(stamp);
    msg->imuAvailable = imu_available; msg->odomAvailable = odom_available;
    if (rpy_init) { msg->imuRollInit = rpy_init[0]; msg->imuPitchInit = rpy_init[1]; msg->imuYawInit = rpy_init[2]; }
    if (guess) { msg->initialGuessX = guess[0]; msg->initialGuessY = guess[1]; msg->initialGuessZ = guess[2]; msg->initialGuessRoll = guess[3]; msg->initialGuessPitch = guess[4]; msg->initialGuessYaw = guess[5]; }
    pcl::PointCloud<PointType> c; fill_cloud(xyzi, n, c);
    pcl::toROSMsg(c, msg->cloud_deskewed);
    mo->laserCloudInfoHandler(msg);
}

// state after a frame: transformTobeMapped (roll, pitch, yaw, x, y, z), counts[0..4] = keyframes, laserCloudSurfLastDSNum, laserCloudSurfFromMapDSNum, isDegenerate, ScanContext entries
void refmo_state(void* h, float* tf6, int* counts) {
    mapOptimization* mo = (mapOptimization*)h;
    std::memcpy(tf6, mo->transformTobeMapped, 6 * sizeof(float));
    counts[0] = (int)mo->cloudKeyPoses3D->size(); counts[1] = mo->laserCloudSurfLastDSNum; counts[2] = mo->laserCloudSurfFromMapDSNum; counts[3] = mo->isDegenerate ? 1 : 0;
    counts[4] = (int)mo->scManager.polarcontexts_.size();
}
void refmo_set_transform(void* h, const float* tf6) { std::memcpy(((mapOptimization*)h)->transformTobeMapped, tf6, 6 * sizeof(float)); }
// which: 0 laserCloudSurfLastDS, 1 laserCloudSurfFromMapDS, 2 laserCloudSurfFromMap (before the filter), 3 laserCloudSurfLast, 100 + k = surfCloudKeyFrames[k]; returns the size
int refmo_get_cloud(void* h, int which, float* out, int cap) {
    mapOptimization* mo = (mapOptimization*)h;
    if (which >= 100) return which - 100 < (int)mo->surfCloudKeyFrames.size() ? copy_cloud(*mo->surfCloudKeyFrames[which - 100], out, cap) : -1;
    return copy_cloud(which == 0 ? *mo->laserCloudSurfLastDS : which == 1 ? *mo->laserCloudSurfFromMapDS : which == 2 ? *mo->laserCloudSurfFromMap : *mo->laserCloudSurfLast, out, cap);
}
int refmo_get_keypose(void* h, int k, float* pose6, double* time) {
    mapOptimization* mo = (mapOptimization*)h;
    if (k < 0 || k >= (int)mo->cloudKeyPoses6D->size()) return -1;
    const PointTypePose& p = mo->cloudKeyPoses6D->points[k];
    pose6[0] = p.roll; pose6[1] = p.pitch; pose6[2] = p.yaw; pose6[3] = p.x; pose6[4] = p.y; pose6[5] = p.z; *time = p.time;
    return 0;
}

// ---- function-level entries (the members are public in the reference) ----
// loads laserCloudSurfLastDS and laserCloudSurfFromMapDS directly and builds the kd-tree the way scan2MapOptimization does (:1302)
void refmo_set_scan_and_map(void* h, const float* scan_ds, int n, const float* map_ds, int m) {
    mapOptimization* mo = (mapOptimization*)h;
    fill_cloud(scan_ds, n, *mo->laserCloudSurfLastDS); mo->laserCloudSurfLastDSNum = n;
    fill_cloud(map_ds, m, *mo->laserCloudSurfFromMapDS); mo->laserCloudSurfFromMapDSNum = m;
    mo->kdtreeSurfFromMap->setInputCloud(mo->laserCloudSurfFromMapDS);
    if ((int)mo->laserCloudOriSurfVec.size() < n) { mo->laserCloudOriSurfVec.resize(n); mo->coeffSelSurfVec.resize(n); mo->laserCloudOriSurfFlag.assign(n, false); }
}
// the body of the loop at :1304-1314 for one iterCount: returns LMOptimization's result; n_sel = laserCloudOri->size()
int refmo_iteration(void* h, int iterCount, float* tf6_out, int* n_sel) {
    mapOptimization* mo = (mapOptimization*)h;
    mo->laserCloudOri->clear();
    mo->coeffSel->clear();
    mo->surfOptimization();
    mo->combineOptimizationCoeffs();
    const bool conv = mo->LMOptimization(iterCount);
    if (n_sel) *n_sel = (int)mo->laserCloudOri->size();
    if (tf6_out) std::memcpy(tf6_out, mo->transformTobeMapped, 6 * sizeof(float));
    return conv ? 1 : 0;
}
// surfOptimization alone (:1074-1143): per-point coeff (x, y, z, intensity) and flag for the current transformTobeMapped
void refmo_surf_optimization(void* h, float* coeff4, unsigned char* flag) {
    mapOptimization* mo = (mapOptimization*)h;
    std::fill(mo->laserCloudOriSurfFlag.begin(), mo->laserCloudOriSurfFlag.end(), false);
    mo->surfOptimization();
    for (int i = 0; i < mo->laserCloudSurfLastDSNum; ++i) {
        flag[i] = mo->laserCloudOriSurfFlag[i] ? 1 : 0;
        const PointType& c = mo->coeffSelSurfVec[i];
        coeff4[4 * i] = flag[i] ? c.x : 0.f; coeff4[4 * i + 1] = flag[i] ? c.y : 0.f; coeff4[4 * i + 2] = flag[i] ? c.z : 0.f; coeff4[4 * i + 3] = flag[i] ? c.intensity : 0.f;
    }
    std::fill(mo->laserCloudOriSurfFlag.begin(), mo->laserCloudOriSurfFlag.end(), false);
}
// scan2MapOptimization() as a whole (:1295-1321) on the loaded scan / map; needs at least one key pose (:1297) — the harness adds a dummy one if there is none
void refmo_scan2map(void* h, float* tf6_inout, int imu_available, const float* rpy_init) {
    mapOptimization* mo = (mapOptimization*)h;
    if (mo->cloudKeyPoses3D->points.empty()) { PointType p; mo->cloudKeyPoses3D->push_back(p); }
    mo->cloudInfo.imuAvailable = imu_available;
    if (rpy_init) { mo->cloudInfo.imuRollInit = rpy_init[0]; mo->cloudInfo.imuPitchInit = rpy_init[1]; mo->cloudInfo.imuYawInit = rpy_init[2]; }
    std::memcpy(mo->transformTobeMapped, tf6_inout, 6 * sizeof(float));
    // (scan2MapOptimization rebuilds the kd-tree itself)
    mo->scan2MapOptimization();
    std::memcpy(tf6_inout, mo->transformTobeMapped, 6 * sizeof(float));
}
int refmo_get_lm_state(void* h, float* matP36) {
    mapOptimization* mo = (mapOptimization*)h;
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) matP36[i * 6 + j] = mo->matP.at<float>(i, j);
    return mo->isDegenerate ? 1 : 0;
}
// ---- the scalar member functions one by one (fuzzed against the library's host entries) ----
static void set_cloud_info(mapOptimization* mo, const float* ci11) {   // (imuAvailable, odomAvailable, imuRollInit, imuPitchInit, imuYawInit, initialGuessX, Y, Z, Roll, Pitch, Yaw)
    mo->cloudInfo.imuAvailable = (int64_t)ci11[0]; mo->cloudInfo.odomAvailable = (int64_t)ci11[1];
    mo->cloudInfo.imuRollInit = ci11[2]; mo->cloudInfo.imuPitchInit = ci11[3]; mo->cloudInfo.imuYawInit = ci11[4];
    mo->cloudInfo.initialGuessX = ci11[5]; mo->cloudInfo.initialGuessY = ci11[6]; mo->cloudInfo.initialGuessZ = ci11[7];
    mo->cloudInfo.initialGuessRoll = ci11[8]; mo->cloudInfo.initialGuessPitch = ci11[9]; mo->cloudInfo.initialGuessYaw = ci11[10];
}
// updateInitialGuess() (:899-958); have_keyframes = whether cloudKeyPoses3D is non-empty (the harness adds / removes one dummy key pose)
void refmo_update_initial_guess(void* h, const float* ci11, int have_keyframes, float* tf6_inout) {
    mapOptimization* mo = (mapOptimization*)h;
    set_cloud_info(mo, ci11);
    if (have_keyframes && mo->cloudKeyPoses3D->points.empty()) { PointType p; mo->cloudKeyPoses3D->push_back(p); }
    if (!have_keyframes) mo->cloudKeyPoses3D->clear();
    std::memcpy(mo->transformTobeMapped, tf6_inout, 6 * sizeof(float));
    mo->updateInitialGuess();
    std::memcpy(tf6_inout, mo->transformTobeMapped, 6 * sizeof(float));
}
// transformUpdate() (:1323-1353)
void refmo_transform_update(void* h, const float* ci11, float* tf6_inout) {
    mapOptimization* mo = (mapOptimization*)h;
    set_cloud_info(mo, ci11);
    std::memcpy(mo->transformTobeMapped, tf6_inout, 6 * sizeof(float));
    mo->transformUpdate();
    std::memcpy(tf6_inout, mo->transformTobeMapped, 6 * sizeof(float));
}
// saveFrame() (:1365-1384) against a last key pose (NULL = no keyframes yet)
int refmo_save_frame(void* h, const float* last_pose6, const float* tf6) {
    mapOptimization* mo = (mapOptimization*)h;
    mo->cloudKeyPoses3D->clear(); mo->cloudKeyPoses6D->clear();
    if (last_pose6) {
        PointType p3; PointTypePose p6;
        p3.x = last_pose6[3]; p3.y = last_pose6[4]; p3.z = last_pose6[5];
        p6.x = p3.x; p6.y = p3.y; p6.z = p3.z; p6.roll = last_pose6[0]; p6.pitch = last_pose6[1]; p6.yaw = last_pose6[2]; p6.time = 0; p6.intensity = 0;
        mo->cloudKeyPoses3D->push_back(p3); mo->cloudKeyPoses6D->push_back(p6);
    }
    std::memcpy(mo->transformTobeMapped, tf6, 6 * sizeof(float));
    return mo->saveFrame() ? 1 : 0;
}

// LMOptimization(iterCount) (:1158-1293) on given correspondences: laserCloudOri / coeffSel are filled directly (they are what combineOptimizationCoeffs leaves)
int refmo_lm_optimization(void* h, int iterCount, const float* ori4, const float* coeff4, int n, float* tf6_inout) {
    mapOptimization* mo = (mapOptimization*)h;
    fill_cloud(ori4, n, *mo->laserCloudOri); fill_cloud(coeff4, n, *mo->coeffSel);
    std::memcpy(mo->transformTobeMapped, tf6_inout, 6 * sizeof(float));
    const bool conv = mo->LMOptimization(iterCount);
    std::memcpy(tf6_inout, mo->transformTobeMapped, 6 * sizeof(float));
    return conv ? 1 : 0;
}

// ---- bench.py --impl reference / cpu_baseline: one headline step timed on the reference's own member functions ----
// laserCloudSurfLast = scan (filled before the clock starts, like the GPU arm's resident scan), then downsampleCurrentScan() (:1061-1067), the kd-tree
// build of scan2MapOptimization (:1302) and `iters` passes of its loop body (:1306-1314) — the convergence break is left out when force_all != 0, as the
// headline configuration asks on both sides.  timings_ms = {downsample, kd-tree build, surfOptimization + combine, LMOptimization}.  Returns the iterations run.
int refmo_bench_step(void* h, const float* scan_xyzi, int n, const float* tf6_init, int iters, int force_all, float* tf6_out, double* timings_ms) {
    mapOptimization* mo = (mapOptimization*)h;
    using clk = std::chrono::steady_clock;
    auto ms = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    fill_cloud(scan_xyzi, n, *mo->laserCloudSurfLast);
    std::memcpy(mo->transformTobeMapped, tf6_init, 6 * sizeof(float));
    std::fill(mo->laserCloudOriSurfFlag.begin(), mo->laserCloudOriSurfFlag.end(), false);
    auto t0 = clk::now();
    mo->downsampleCurrentScan();
    auto t1 = clk::now();
    if ((int)mo->laserCloudOriSurfVec.size() < mo->laserCloudSurfLastDSNum) {
        mo->laserCloudOriSurfVec.resize(mo->laserCloudSurfLastDSNum); mo->coeffSelSurfVec.resize(mo->laserCloudSurfLastDSNum); mo->laserCloudOriSurfFlag.assign(mo->laserCloudSurfLastDSNum, false);
    }
    mo->kdtreeSurfFromMap->setInputCloud(mo->laserCloudSurfFromMapDS);
    auto t2 = clk::now();
    double t_surf = 0, t_lm = 0; int it = 0;
    for (; it < iters; ++it) {
        auto a = clk::now();
        mo->laserCloudOri->clear();
        mo->coeffSel->clear();
        mo->surfOptimization();
        mo->combineOptimizationCoeffs();
        auto b = clk::now();
        const bool conv = mo->LMOptimization(it);
        auto c = clk::now();
        t_surf += ms(a, b); t_lm += ms(b, c);
        if (conv && !force_all) { ++it; break; }
    }
    if (tf6_out) std::memcpy(tf6_out, mo->transformTobeMapped, 6 * sizeof(float));
    if (timings_ms) { timings_ms[0] = ms(t0, t1); timings_ms[1] = ms(t1, t2); timings_ms[2] = t_surf; timings_ms[3] = t_lm; }
    return it;
}
void refmo_set_map(void* h, const float* map_ds, int m) {
    mapOptimization* mo = (mapOptimization*)h;
    fill_cloud(map_ds, m, *mo->laserCloudSurfFromMapDS); mo->laserCloudSurfFromMapDSNum = m;
}

// ---- the rows next to the path ----
// harness set-up: appends a keyframe exactly as saveKeyFramesAndFactor stores one (:1548-1580: cloudKeyPoses3D / 6D with intensity = index, surfCloudKeyFrames)
void refmo_add_keyframe(void* h, const float* xyzi, int n, const float* pose6, double time) {
    mapOptimization* mo = (mapOptimization*)h;
    PointType p3; PointTypePose p6;
    p3.x = pose6[3]; p3.y = pose6[4]; p3.z = pose6[5]; p3.intensity = mo->cloudKeyPoses3D->size();
    p6.x = p3.x; p6.y = p3.y; p6.z = p3.z; p6.intensity = p3.intensity; p6.roll = pose6[0]; p6.pitch = pose6[1]; p6.yaw = pose6[2]; p6.time = time;
    mo->cloudKeyPoses3D->push_back(p3); mo->cloudKeyPoses6D->push_back(p6);
    pcl::PointCloud<PointType>::Ptr c(new pcl::PointCloud<PointType>()); fill_cloud(xyzi, n, *c);
    mo->surfCloudKeyFrames.push_back(c);
    mo->timeLaserInfoCur = time;
}
// extractSurroundingKeyFrames() (:1046-1059 → extractNearby :975-1010 → extractCloud :1012-1044) at time t; laserCloudSurfFromMap (get_cloud 2) then holds the
// selected keyframes' transformed clouds in selection order
void refmo_extract_surrounding_keyframes(void* h, double t) {
    mapOptimization* mo = (mapOptimization*)h;
    mo->timeLaserInfoCur = t;
    mo->extractSurroundingKeyFrames();
}
// loopFindNearKeyframes (:821-844) on a snapshot of the key poses taken the way performSCLoopClosure takes it (:629-632); returns the size
int refmo_loop_find_near_keyframes(void* h, int key, int search_num, int loop_index, float* out, int cap) {
    mapOptimization* mo = (mapOptimization*)h;
    *mo->copy_cloudKeyPoses3D = *mo->cloudKeyPoses3D;
    *mo->copy_cloudKeyPoses6D = *mo->cloudKeyPoses6D;
    pcl::PointCloud<PointType>::Ptr c(new pcl::PointCloud<PointType>());
    mo->loopFindNearKeyframes(c, key, search_num, loop_index);
    return copy_cloud(*c, out, cap);
}
// publishGlobalMap (:453-502) with one subscriber on liorf/mapping/map_global: the published cloud; returns its size (-1: nothing was published)
int refmo_publish_global_map(void* h, float* out, int cap) {
    mapOptimization* mo = (mapOptimization*)h;
    const std::string topic = "liorf/mapping/map_global";
    ros::shim::subscribers()[topic] = 1;
    const long before = ros::shim::outcount<sensor_msgs::PointCloud2>()[topic];
    mo->publishGlobalMap();
    ros::shim::subscribers()[topic] = 0;
    if (ros::shim::outcount<sensor_msgs::PointCloud2>()[topic] == before) return -1;
    pcl::PointCloud<PointType> c; pcl::fromROSMsg(ros::shim::outbox<sensor_msgs::PointCloud2>()[topic], c);
    return copy_cloud(c, out, cap);
}

// detectLoopClosureID of the node's SCManager (performSCLoopClosure :636)
void refmo_sc_detect(void* h, int* loop_id, float* yaw) { auto r = ((mapOptimization*)h)->scManager.detectLoopClosureID(); *loop_id = r.first; *yaw = r.second; }

}  // extern "C"

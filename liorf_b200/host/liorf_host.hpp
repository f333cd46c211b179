// liorf_host.hpp — C++ host-side mirror of the reference's member-function interface for the hot path, on top of the
// C ABI (include/liorf_b200.h).  Same names, argument meaning and (silent early-return) error behaviour as
//   ImageProjection::projectPointCloud                     src/imageProjection.cpp:568
//   mapOptimization::{downsampleCurrentScan, extractSurroundingKeyFrames, scan2MapOptimization, surfOptimization,
//                     combineOptimizationCoeffs, LMOptimization, saveFrame}   src/mapOptmization.cpp:1061,1046,1295,1074,1145,1158,1365
//   SCManager::{makeAndSaveScancontextAndKeys, detectLoopClosureID}           include/Scancontext.h:72-73
// The implicit member state of the reference classes (laserCloudSurfLast, laserCloudSurfLastDS, laserCloudSurfFromMapDS,
// transformTobeMapped, isDegenerate/matP, the ScanContext database ...) lives in the liorf_ctx on the device; the
// members kept here are the few the surrounding ROS code reads.  No ROS, PCL, GTSAM or OpenCV.
#pragma once
#include <cstdio>
#include <stdexcept>
#include <utility>
#include <vector>
#include "../../include/liorf_b200.h"

namespace liorf_b200 {

struct ParamServer {                                   // include/utility.h:68-291, the subset the path reads
    liorf_params p;
    float surroundingkeyframeAddingDistThreshold = 1.0f, surroundingkeyframeAddingAngleThreshold = 0.2f;
    float surroundingKeyframeDensity = 2.0f, z_tollerance = 1000.f, rotation_tollerance = 1000.f;
    int imuType = 0, useImuHeadingInitialization = 1; float imuRPYWeight = 0.01f;      // utility.h: imuType, useImuHeadingInitialization, imuRPYWeight
    float historyKeyframeSearchRadius = 10.0f, historyKeyframeFitnessScore = 0.3f, loopClosureICPSurfLeafSize = 0.5f;   // utility.h:239-248
    int historyKeyframeSearchNum = 25;
    ParamServer() { liorf_default_params(&p); }
};

class Context {
public:
    explicit Context(const ParamServer& ps) {
        if (liorf_create(&ps.p, &ctx_) != 0) throw std::runtime_error("liorf_create failed (this library has no CPU path)");
    }
    ~Context() { liorf_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    liorf_ctx* get() const { return ctx_; }
private:
    liorf_ctx* ctx_ = nullptr;
};

class ImageProjection {
public:
    ImageProjection(Context& c) : ctx(c) {}
    // members the reference's imuDeskewInfo fills (src/imageProjection.cpp:86-89, 350-409)
    std::vector<double> imuTime, imuRotX, imuRotY, imuRotZ;
    int imuPointerCur = -1;
    double timeScanCur = 0.0;
    int deskewFlag = 1; bool imuAvailable = false;
    std::vector<liorf_point_xyzirt> laserCloudIn;
    std::vector<liorf_point> fullCloud;                // filled only when copyOut is set
    bool copyOut = false;

    void projectPointCloud() {
        const int deskew = !(deskewFlag == -1 || !imuAvailable);                 // :538
        int n_out = 0;
        if (copyOut) fullCloud.resize(laserCloudIn.size() ? laserCloudIn.size() : 1);
        liorf_project_point_cloud(ctx.get(), laserCloudIn.data(), (int)laserCloudIn.size(), timeScanCur, imuTime.data(), imuRotX.data(), imuRotY.data(),
                                  imuRotZ.data(), imuPointerCur, deskew, copyOut ? fullCloud.data() : nullptr, copyOut ? &n_out : nullptr, nullptr);
        if (copyOut) fullCloud.resize(n_out);
    }
private:
    Context& ctx;
};

class mapOptimization {
public:
    mapOptimization(Context& c, const ParamServer& ps) : ctx(c), P(ps) {}
    float transformTobeMapped[6] = {0, 0, 0, 0, 0, 0};                           // src/mapOptmization.cpp:134
    double timeLaserInfoCur = 0.0;
    bool isDegenerate = false;
    int laserCloudSurfLastDSNum = -1, laserCloudSurfFromMapDSNum = -1;
    liorf_lm_trace lastTrace{};

    void setLaserCloudSurfLast(const std::vector<liorf_point>& cloud) { liorf_set_current_scan(ctx.get(), cloud.data(), (int)cloud.size()); }
    void downsampleCurrentScan() { liorf_downsample_current_scan(ctx.get(), nullptr, nullptr, nullptr); laserCloudSurfLastDSNum = -1; }   // :1061
    void extractSurroundingKeyFrames() {                                        // :1046
        if (liorf_num_keyframes(ctx.get()) <= 0) return;                         // cloudKeyPoses3D->points.empty()
        int ids[4096], n = 0;
        if (liorf_extract_nearby(ctx.get(), timeLaserInfoCur, P.surroundingKeyframeDensity, ids, 4096, &n) != 0) return;
        liorf_extract_surrounding_keyframes(ctx.get(), ids, n, nullptr);
    }
    void scan2MapOptimization() {                                               // :1295
        if (liorf_num_keyframes(ctx.get()) <= 0) return;
        if (liorf_scan2map_optimization(ctx.get(), transformTobeMapped, 30, 0, &lastTrace) != 0) return;
        isDegenerate = lastTrace.degenerate != 0;
        if (!lastTrace.ran) { std::fprintf(stderr, "Not enough features!\n"); return; }   // ROS_WARN at :1319
        transformUpdate();
    }
    liorf_cloud_info_guess cloudInfo{};                                          // the cloud_info fields the two scalar steps read
    void updateInitialGuess() {                                                  // :899-958
        liorf_host_update_initial_guess(&guessState, liorf_num_keyframes(ctx.get()) <= 0, &cloudInfo, P.useImuHeadingInitialization, P.imuType, transformTobeMapped);
    }
    void transformUpdate() {                                                     // :1323-1353
        liorf_host_transform_update(transformTobeMapped, cloudInfo.imuAvailable, P.imuType, cloudInfo.imuRollInit, cloudInfo.imuPitchInit, P.imuRPYWeight,
                                    P.rotation_tollerance, P.z_tollerance);
    }
    bool saveFrame() { return liorf_save_frame(ctx.get(), transformTobeMapped, P.surroundingkeyframeAddingDistThreshold, P.surroundingkeyframeAddingAngleThreshold) == 1; }
    // the part of saveKeyFramesAndFactor that touches the hot path's data (:1576-1595); the factor graph stays outside
    int saveKeyFrame() { int id = liorf_add_keyframe(ctx.get(), transformTobeMapped, timeLaserInfoCur); if (id >= 0) liorf_sc_make_and_save(ctx.get(), nullptr, 0); return id; }
    // performSCLoopClosure (:624-730) up to the point where the reference queues the loop factor: detectLoopClosureID, the two
    // clouds of loopFindNearKeyframes (base_key = 0), pcl::IterativeClosestPoint, the fitness gate.  True = a factor
    // (loopKeyCur, loopKeyPre, icp.pose6) would be pushed to loopIndexQueue / loopPoseQueue.
    bool performSCLoopClosure(int& loopKeyCur, int& loopKeyPre, liorf_icp_result& icp) {
        const int n = liorf_num_keyframes(ctx.get());
        if (n <= 0) return false;                                                 // :626
        float yawDiffRad = 0.f;
        loopKeyCur = n - 1; loopKeyPre = -1;
        if (liorf_sc_detect_loop_closure_id(ctx.get(), &loopKeyPre, &yawDiffRad, nullptr, nullptr) != 0 || loopKeyPre == -1) return false;   // :636-641
        for (const auto& e : loopIndexContainer) if (e.first == loopKeyCur) return false;                                                  // :643-645
        if (liorf_loop_closure_icp(ctx.get(), loopKeyCur, loopKeyPre, P.historyKeyframeSearchNum, 0, P.loopClosureICPSurfLeafSize,
                                   P.historyKeyframeSearchRadius * 2, 100, &icp) != 0 || !icp.ran) return false;                            // :652-671
        if (!icp.converged || icp.fitness > P.historyKeyframeFitnessScore) return false;                                                    // :673
        loopIndexContainer.emplace_back(loopKeyCur, loopKeyPre);                                                                            // :727
        return true;
    }
    std::vector<std::pair<int, int>> loopIndexContainer;
    // per-function forms used by the parity tests
    void surfOptimization(std::vector<liorf_point>& coeffSelSurfVec, std::vector<uint8_t>& laserCloudOriSurfFlag) {            // :1074
        int n = 0; liorf_get_scan_ds(ctx.get(), nullptr, 0, &n); laserCloudSurfLastDSNum = n;
        coeffSelSurfVec.resize(n ? n : 1); laserCloudOriSurfFlag.resize(n ? n : 1);
        liorf_surf_optimization(ctx.get(), transformTobeMapped, coeffSelSurfVec.data(), laserCloudOriSurfFlag.data(), nullptr, nullptr, nullptr, nullptr);
        coeffSelSurfVec.resize(n); laserCloudOriSurfFlag.resize(n);
    }
    int combineOptimizationCoeffs() { int n = 0; liorf_combine_optimization_coeffs(ctx.get(), nullptr, nullptr, &n); return n; }   // :1145
    bool LMOptimization(int iterCount) { return liorf_lm_optimization(ctx.get(), iterCount, transformTobeMapped, nullptr, nullptr, nullptr, nullptr) == 1; }   // :1158
private:
    Context& ctx;
    const ParamServer& P;
    liorf_guess_state guessState{};
};

class SCManager {
public:
    explicit SCManager(Context& c) : ctx(c) {}
    void makeAndSaveScancontextAndKeys(const std::vector<liorf_point>& scan_down) { liorf_sc_make_and_save(ctx.get(), scan_down.data(), (int)scan_down.size()); }
    std::pair<int, float> detectLoopClosureID() {
        int id = -1; float yaw = 0.f;
        liorf_sc_detect_loop_closure_id(ctx.get(), &id, &yaw, nullptr, nullptr);
        return {id, yaw};
    }
private:
    Context& ctx;
};

}  // namespace liorf_b200

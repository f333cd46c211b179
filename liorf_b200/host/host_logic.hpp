// host_logic.hpp — the small scalar host-side steps around the GPU hot path, written as plain C++ (no CUDA, no ROS,
// no PCL).  These mirror reference code that is host logic in liorf too:
//   extract_nearby            mapOptimization::extractNearby            src/mapOptmization.cpp:975-1010 (keyframe SELECTION only)
//   save_frame                mapOptimization::saveFrame                src/mapOptmization.cpp:1365-1384
//   transform_update_clamp    mapOptimization::transformUpdate          src/mapOptmization.cpp:1348-1350 (the clamps)
//   transform_update          mapOptimization::transformUpdate          src/mapOptmization.cpp:1323-1353 (9-axis roll/pitch slerp + clamps)
//   update_initial_guess      mapOptimization::updateInitialGuess       src/mapOptmization.cpp:899-958
//   icp_estimate / IcpConvergence   the scalar half of pcl::IterativeClosestPoint as performSCLoopClosure configures it (:652-674)
// They run on a few hundred key poses per frame; the heavy half of extractSurroundingKeyFrames (extractCloud) is CUDA.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <utility>
#include <vector>

// the ICP scalar half (sym3_jacobi .. IcpConvergence) also runs on the device: the last block of an ICP iteration calls it (csrc/icp.cuh)
#ifdef __CUDACC__
#define LIORF_HD __host__ __device__
#else
#define LIORF_HD
#endif

namespace liorf_host {

struct KeyPose { float roll, pitch, yaw, x, y, z; double time; };

inline void get_transformation(float x, float y, float z, float roll, float pitch, float yaw, float t[12]) {   // pcl::getTransformation
    float A = std::cos(yaw), B = std::sin(yaw), C = std::cos(pitch), D = std::sin(pitch);
    float E = std::cos(roll), F = std::sin(roll), DE = D * E, DF = D * F;
    t[0] = A * C;  t[1] = A * DF - B * E;  t[2]  = B * F + A * DE;  t[3]  = x;
    t[4] = B * C;  t[5] = A * E + B * DF;  t[6]  = B * DE - A * F;  t[7]  = y;
    t[8] = -D;     t[9] = C * F;           t[10] = C * E;           t[11] = z;
}
inline void affine_inverse(const float t[12], float o[12]) {          // Eigen::Affine3f::inverse(): 3x3 cofactors, -(L^-1 t)
    auto M = [&](int r, int c) { return t[r * 4 + c]; };
    auto cof = [&](int i, int j) { int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3; return M(i1, j1) * M(i2, j2) - M(i1, j2) * M(i2, j1); };
    float c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
    float det = (c00 * M(0, 0) + c10 * M(1, 0)) + c20 * M(2, 0), invdet = 1.f / det;
    float L[3][3] = {{c00 * invdet, c10 * invdet, c20 * invdet},
                     {cof(0, 1) * invdet, cof(1, 1) * invdet, cof(2, 1) * invdet},
                     {cof(0, 2) * invdet, cof(1, 2) * invdet, cof(2, 2) * invdet}};
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) o[r * 4 + c] = L[r][c]; o[r * 4 + 3] = -((L[r][0] * t[3] + L[r][1] * t[7]) + L[r][2] * t[11]); }
}
inline void affine_mul(const float a[12], const float b[12], float o[12]) {
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) o[r * 4 + c] = (a[r * 4] * b[c] + a[r * 4 + 1] * b[4 + c]) + a[r * 4 + 2] * b[8 + c];
        o[r * 4 + 3] = ((a[r * 4] * b[3] + a[r * 4 + 1] * b[7]) + a[r * 4 + 2] * b[11]) + a[r * 4 + 3];
    }
}
inline void get_translation_and_euler(const float t[12], float& x, float& y, float& z, float& roll, float& pitch, float& yaw) {   // pcl
    x = t[3]; y = t[7]; z = t[11];
    roll = std::atan2(t[9], t[10]); pitch = std::asin(-t[8]); yaw = std::atan2(t[4], t[0]);
}

// saveFrame (:1365-1384).  tf = transformTobeMapped (roll,pitch,yaw,x,y,z).
inline bool save_frame(const KeyPose* last /*nullptr when no keyframe yet*/, const float tf[6], float dist_thr, float ang_thr) {
    if (!last) return true;
    float a[12], b[12], ai[12], bt[12];
    get_transformation(last->x, last->y, last->z, last->roll, last->pitch, last->yaw, a);
    get_transformation(tf[3], tf[4], tf[5], tf[0], tf[1], tf[2], b);
    affine_inverse(a, ai); affine_mul(ai, b, bt);
    float x, y, z, roll, pitch, yaw; get_translation_and_euler(bt, x, y, z, roll, pitch, yaw);
    if (std::fabs(roll) < ang_thr && std::fabs(pitch) < ang_thr && std::fabs(yaw) < ang_thr && std::sqrt(x * x + y * y + z * z) < dist_thr) return false;
    return true;
}

inline float constraint_transformation(float value, float limit) { if (value < -limit) value = -limit; if (value > limit) value = limit; return value; }
// ImageProjection::deskewInfo's IMU part (src/imageProjection.cpp:330-409) on a caller-held sample array instead of the ROS deque:
// stamps ascending, angular velocities already rotated into the lidar frame (imuConverter, include/utility.h:257-272).
//   gate (:337): no sample, first sample after timeScanCur or last sample before timeScanEnd → "waiting for IMU data" (returns -1 rows)
//   pop  (:354-360): samples older than timeScanCur - 0.01 leave the queue → n_pop
//   table(:367-401): row 0 = zeros at the first remaining stamp, then rectangle-rule integration rot[k] = rot[k-1] + w[k] (t[k] - t[k-1])
//                    with the CURRENT sample's rate, until a stamp exceeds timeScanEnd + 0.01
//   rpy  (:371-375): index of the last sample at or before timeScanCur whose orientation gives imuRoll/Pitch/YawInit (imuType != 0)
// imuPointerCur = rows - 1; imuAvailable iff imuPointerCur > 0 (:403-408).  capacity = queueLength (2000, :62): the reference does not
// check it; here a longer table is an error (-2).
struct ImuDeskewInfo { int imu_pointer_cur = -1; bool imu_available = false; int n_pop = 0; int rpy_index = -1; };
inline int imu_deskew_info(const double* stamp, const double* gyro_xyz, int n, double timeScanCur, double timeScanEnd, bool check_gate,
                           double* imuTime, double* imuRotX, double* imuRotY, double* imuRotZ, int capacity, ImuDeskewInfo& out) {
    out = ImuDeskewInfo();
    if (check_gate && (n == 0 || stamp[0] > timeScanCur || stamp[n - 1] < timeScanEnd)) return -1;
    int first = 0;
    while (first < n && stamp[first] < timeScanCur - 0.01) ++first;
    out.n_pop = first;
    if (first == n) return 0;
    int ptr = 0;
    for (int i = first; i < n; ++i) {
        const double t = stamp[i];
        if (t <= timeScanCur) out.rpy_index = i;
        if (t > timeScanEnd + 0.01) break;
        if (ptr >= capacity) return -2;
        if (ptr == 0) { imuRotX[0] = 0; imuRotY[0] = 0; imuRotZ[0] = 0; imuTime[0] = t; ++ptr; continue; }
        const double timeDiff = t - imuTime[ptr - 1];
        imuRotX[ptr] = imuRotX[ptr - 1] + gyro_xyz[3 * i] * timeDiff;
        imuRotY[ptr] = imuRotY[ptr - 1] + gyro_xyz[3 * i + 1] * timeDiff;
        imuRotZ[ptr] = imuRotZ[ptr - 1] + gyro_xyz[3 * i + 2] * timeDiff;
        imuTime[ptr] = t;
        ++ptr;
    }
    out.imu_pointer_cur = ptr - 1;
    out.imu_available = out.imu_pointer_cur > 0;
    return ptr;
}

inline void transform_update_clamp(float tf[6], float rotation_tollerance, float z_tollerance) {                // :1348-1350
    tf[0] = constraint_transformation(tf[0], rotation_tollerance);
    tf[1] = constraint_transformation(tf[1], rotation_tollerance);
    tf[5] = constraint_transformation(tf[5], z_tollerance);
}


// ---- transformUpdate with the 9-axis IMU fusion (:1323-1353) --------------------------------------------------------
// tf::Quaternion / tf::Matrix3x3 (tfScalar = double) restated for the two single-axis blends the reference performs.
struct Quat { double x, y, z, w; };
inline Quat quat_set_rpy(double roll, double pitch, double yaw) {                 // tf::Quaternion::setRPY
    const double hy = yaw * 0.5, hp = pitch * 0.5, hr = roll * 0.5;
    const double cy = std::cos(hy), sy = std::sin(hy), cp = std::cos(hp), sp = std::sin(hp), cr = std::cos(hr), sr = std::sin(hr);
    return Quat{sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy};
}
inline double quat_dot(const Quat& a, const Quat& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline Quat quat_slerp(const Quat& a, const Quat& q, double t) {                  // tf::Quaternion::slerp (shortest path)
    const double s = std::sqrt(quat_dot(a, a) * quat_dot(q, q));
    const double dp = quat_dot(a, q);
    const double theta = (dp < 0 ? std::acos(-dp / s) * 2.0 : std::acos(dp / s) * 2.0) / 2.0;       // angleShortestPath(q) / 2
    if (theta != 0.0) {
        const double d = 1.0 / std::sin(theta), s0 = std::sin((1.0 - t) * theta), s1 = std::sin(t * theta);
        if (dp < 0) return Quat{(a.x * s0 + -q.x * s1) * d, (a.y * s0 + -q.y * s1) * d, (a.z * s0 + -q.z * s1) * d, (a.w * s0 + -q.w * s1) * d};
        return Quat{(a.x * s0 + q.x * s1) * d, (a.y * s0 + q.y * s1) * d, (a.z * s0 + q.z * s1) * d, (a.w * s0 + q.w * s1) * d};
    }
    return a;
}
inline void quat_get_rpy(const Quat& q, double& roll, double& pitch, double& yaw) {   // tf::Matrix3x3(q).getRPY → getEulerYPR, solution 1
    const double d = quat_dot(q, q), s = 2.0 / d;
    const double xs = q.x * s, ys = q.y * s, zs = q.z * s;
    const double wx = q.w * xs, wy = q.w * ys, wz = q.w * zs, xx = q.x * xs, xy = q.x * ys, xz = q.x * zs, yy = q.y * ys, yz = q.y * zs, zz = q.z * zs;
    const double m00 = 1.0 - (yy + zz), m10 = xy + wz, m20 = xz - wy, m21 = yz + wx, m22 = 1.0 - (xx + yy);
    if (std::fabs(m20) >= 1.0) {                                                 // gimbal lock branch of getEulerYPR
        yaw = 0.0;
        const double delta = std::atan2(m21, m22);
        if (m20 < 0) { pitch = M_PI / 2.0; roll = delta; } else { pitch = -M_PI / 2.0; roll = delta; }
        return;
    }
    pitch = -std::asin(m20);
    roll = std::atan2(m21 / std::cos(pitch), m22 / std::cos(pitch));
    yaw = std::atan2(m10 / std::cos(pitch), m00 / std::cos(pitch));
}
inline void transform_update(float tf[6], bool imu_available, int imu_type, float imu_roll_init, float imu_pitch_init, float imu_rpy_weight,
                             float rotation_tollerance, float z_tollerance) {
    if (imu_available && imu_type) {                                             // :1325
        if (std::abs(imu_pitch_init) < 1.4) {                                    // :1327 (float abs compared with the double literal)
            const double w = imu_rpy_weight;
            double r, p, y;
            quat_get_rpy(quat_slerp(quat_set_rpy(tf[0], 0, 0), quat_set_rpy(imu_roll_init, 0, 0), w), r, p, y);   // :1335-1338
            tf[0] = (float)r;
            quat_get_rpy(quat_slerp(quat_set_rpy(0, tf[1], 0), quat_set_rpy(0, imu_pitch_init, 0), w), r, p, y);  // :1341-1344
            tf[1] = (float)p;
        }
    }
    transform_update_clamp(tf, rotation_tollerance, z_tollerance);              // :1348-1350
}

// ---- updateInitialGuess (:899-958) -----------------------------------------------------------------------------------
// The function-local statics of the reference become an explicit state object.
struct InitialGuessState {
    float lastImuTransformation[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    bool lastImuPreTransAvailable = false;
    float lastImuPreTransformation[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
};
struct CloudInfoGuess {                      // the cloud_info fields the function reads (msg/cloud_info.msg)
    int imuAvailable, odomAvailable;
    float imuRollInit, imuPitchInit, imuYawInit;
    float initialGuessX, initialGuessY, initialGuessZ, initialGuessRoll, initialGuessPitch, initialGuessYaw;
};
inline void update_initial_guess(InitialGuessState& st, bool no_keyframes_yet, const CloudInfoGuess& ci, bool useImuHeadingInitialization, int imuType, float tf[6]) {
    if (no_keyframes_yet) {                                                      // :906-917
        tf[0] = ci.imuRollInit; tf[1] = ci.imuPitchInit; tf[2] = ci.imuYawInit;
        if (!useImuHeadingInitialization) tf[2] = 0;
        get_transformation(0, 0, 0, ci.imuRollInit, ci.imuPitchInit, ci.imuYawInit, st.lastImuTransformation);
        return;
    }
    if (ci.odomAvailable) {                                                      // :922-943 IMU pre-integration increment
        float transBack[12];
        get_transformation(ci.initialGuessX, ci.initialGuessY, ci.initialGuessZ, ci.initialGuessRoll, ci.initialGuessPitch, ci.initialGuessYaw, transBack);
        if (!st.lastImuPreTransAvailable) {
            for (int i = 0; i < 12; ++i) st.lastImuPreTransformation[i] = transBack[i];
            st.lastImuPreTransAvailable = true;                                  // falls through to the IMU-rotation branch (:945)
        } else {
            float inv[12], incre[12], tobe[12], fin[12];
            affine_inverse(st.lastImuPreTransformation, inv); affine_mul(inv, transBack, incre);
            get_transformation(tf[3], tf[4], tf[5], tf[0], tf[1], tf[2], tobe);
            affine_mul(tobe, incre, fin);
            get_translation_and_euler(fin, tf[3], tf[4], tf[5], tf[0], tf[1], tf[2]);
            for (int i = 0; i < 12; ++i) st.lastImuPreTransformation[i] = transBack[i];
            get_transformation(0, 0, 0, ci.imuRollInit, ci.imuPitchInit, ci.imuYawInit, st.lastImuTransformation);
            return;
        }
    }
    if (ci.imuAvailable && imuType) {                                            // :946-957 rotation-only increment
        float transBack[12], inv[12], incre[12], tobe[12], fin[12];
        get_transformation(0, 0, 0, ci.imuRollInit, ci.imuPitchInit, ci.imuYawInit, transBack);
        affine_inverse(st.lastImuTransformation, inv); affine_mul(inv, transBack, incre);
        get_transformation(tf[3], tf[4], tf[5], tf[0], tf[1], tf[2], tobe);
        affine_mul(tobe, incre, fin);
        get_translation_and_euler(fin, tf[3], tf[4], tf[5], tf[0], tf[1], tf[2]);
        for (int i = 0; i < 12; ++i) st.lastImuTransformation[i] = transBack[i];
    }
}

// ---- scalar half of pcl::IterativeClosestPoint (performSCLoopClosure, :652-674) ------------------------------------------
// TransformationEstimationSVD → Eigen::umeyama without scaling, from the 17 sums the device reduces:
// s[0] = n, s[1..3] = sum of source points, s[4..6] = sum of their target neighbours, s[7..15] = sum d_r * s_c (row r of the
// target point, column c of the source point), s[16] = sum of squared distances.  Returns false for fewer than 3 pairs.
LIORF_HD inline void sym3_jacobi(double A[3][3], double V[3][3]) {      // eigen-decomposition of a symmetric 3x3 by cyclic Jacobi rotations
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) V[i][j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
        const double diag = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2];
        if (off <= 1e-34 * diag || off == 0.0) break;
        for (int p = 0; p < 2; ++p) for (int q = p + 1; q < 3; ++q) {
            if (A[p][q] == 0.0) continue;
            const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
            const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
            const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
            for (int k = 0; k < 3; ++k) { const double akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - sn * akq; A[k][q] = sn * akp + c * akq; }
            for (int k = 0; k < 3; ++k) { const double apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - sn * aqk; A[q][k] = sn * apk + c * aqk; }
            for (int k = 0; k < 3; ++k) { const double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c * vkp - sn * vkq; V[k][q] = sn * vkp + c * vkq; }
        }
    }
}
LIORF_HD inline double det3(const double M[3][3]) {
    return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) + M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
}
LIORF_HD inline bool icp_estimate(const double s[17], float T[16]) {
    const double n = s[0];
    if (n < 3.0) return false;                                   // min_number_correspondences_
    const double ms[3] = {s[1] / n, s[2] / n, s[3] / n}, md[3] = {s[4] / n, s[5] / n, s[6] / n};
    double S[3][3];                                              // sigma = (1/n) dst_demean * src_demean^T
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) S[r][c] = s[7 + 3 * r + c] / n - md[r] * ms[c];
    double A[3][3], V[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[i][j] = S[0][i] * S[0][j] + S[1][i] * S[1][j] + S[2][i] * S[2][j];   // S^T S
    sym3_jacobi(A, V);
    int ord[3] = {0, 1, 2};                                      // singular values descending
    for (int a = 0; a < 2; ++a) for (int b = a + 1; b < 3; ++b) if (A[ord[b]][ord[b]] > A[ord[a]][ord[a]]) { const int t = ord[a]; ord[a] = ord[b]; ord[b] = t; }
    double Vs[3][3], U[3][3], sv[3];
    for (int k = 0; k < 3; ++k) { sv[k] = std::sqrt(A[ord[k]][ord[k]] > 0.0 ? A[ord[k]][ord[k]] : 0.0); for (int i = 0; i < 3; ++i) Vs[i][k] = V[i][ord[k]]; }
    for (int k = 0; k < 2; ++k) {
        double u[3] = {0, 0, 0};
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) u[i] += S[i][j] * Vs[j][k];
        const double nrm = std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
        if (!(nrm > 0.0)) return false;                          // rank < 2: no unique rotation
        for (int i = 0; i < 3; ++i) U[i][k] = u[i] / nrm;
    }
    {   // third left vector: S v3 / s3 when well conditioned, else the completion of the frame (any sign: fixed below)
        double u[3] = {0, 0, 0};
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) u[i] += S[i][j] * Vs[j][2];
        const double nrm = std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
        if (nrm > 1e-9 * sv[0]) { for (int i = 0; i < 3; ++i) U[i][2] = u[i] / nrm; }
        else { U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1]; U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1]; U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1]; }
    }
    const double sgn = det3(U) * det3(Vs) < 0 ? -1.0 : 1.0;     // Eigen::umeyama: S(2) = -1 when det(U) det(V) < 0
    double R[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R[i][j] = U[i][0] * Vs[j][0] + U[i][1] * Vs[j][1] + sgn * U[i][2] * Vs[j][2];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) T[4 * i + j] = (float)R[i][j];
        T[4 * i + 3] = (float)(md[i] - (R[i][0] * ms[0] + R[i][1] * ms[1] + R[i][2] * ms[2]));
    }
    T[12] = 0.f; T[13] = 0.f; T[14] = 0.f; T[15] = 1.f;
    return true;
}
LIORF_HD inline void mat4_mul(const float a[16], const float b[16], float o[16]) {       // Eigen Matrix4f product (float, k ascending)
    float r[16];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { float v = 0.f; for (int k = 0; k < 4; ++k) v += a[4 * i + k] * b[4 * k + j]; r[4 * i + j] = v; }
    for (int k = 0; k < 16; ++k) o[k] = r[k];
}
// pcl::registration::DefaultConvergenceCriteria as pcl::IterativeClosestPoint::computeTransformation configures it (PCL 1.10):
// max iterations, translation threshold = transformation_epsilon (on the SQUARED translation), rotation threshold
// = 1 - transformation_epsilon (on cos of the angle), relative MSE = euclidean_fitness_epsilon, absolute MSE 1e-12,
// max_iterations_similar_transforms = 0.
struct IcpConvergence {
    int max_iterations = 100; double translation_threshold = 1e-6, rotation_threshold = 1.0 - 1e-6, mse_relative = 1e-6, mse_absolute = 1e-12;
    int iterations = 0; double prev_mse = 1.7976931348623157e308;
    enum State { NOT_CONVERGED, ITERATIONS, TRANSFORM, ABS_MSE, REL_MSE, NO_CORRESPONDENCES } state = NOT_CONVERGED;
    LIORF_HD bool has_converged(const float T[16], double cur_mse) {
        state = NOT_CONVERGED;
        ++iterations;
        if (iterations >= max_iterations) { state = ITERATIONS; return true; }
        const double cos_angle = 0.5 * ((double)T[0] + (double)T[5] + (double)T[10] - 1.0);
        const double tr2 = (double)T[3] * T[3] + (double)T[7] * T[7] + (double)T[11] * T[11];
        if (cos_angle >= rotation_threshold && tr2 <= translation_threshold) { state = TRANSFORM; return true; }
        if (std::fabs(cur_mse - prev_mse) < mse_absolute) { state = ABS_MSE; return true; }
        if (std::fabs(cur_mse - prev_mse) / prev_mse < mse_relative) { state = REL_MSE; return true; }
        prev_mse = cur_mse;
        return false;
    }
};

// extractNearby (:975-1010): ids of the keyframes whose clouds extractCloud will fuse, IN ORDER, duplicates included.
// extractCloud's distance gate (:1018) is applied here, on the positions the reference tests: the VOXEL CENTROID for an entry of
// the thinned radius set (its intensity carries the id of the nearest real key pose, which may itself lie beyond the radius),
// the real key pose for an entry of the "younger than 10 s" tail.  include_recent = false gives publishGlobalMap's selection
// (:467-489), which is the same radius / thinning / nearest-1 / gate sequence without the tail.
inline float point_distance(float ax, float ay, float az, float bx, float by, float bz) {   // common_lib::pointDistance(p1, p2), lib/common_lib.cpp:33-37
    return std::sqrt((ax - bx) * (ax - bx) + (ay - by) * (ay - by) + (az - bz) * (az - bz));
}
inline std::vector<int> extract_nearby(const std::vector<KeyPose>& kp, double time_cur, float radius, float density, bool include_recent = true) {
    std::vector<int> out;
    const int n = (int)kp.size();
    if (n == 0) return out;
    const KeyPose& back = kp[n - 1];
    // radiusSearch (FLANN: squared L2_Simple distance strictly below r^2, ascending) — ties by index
    std::vector<std::pair<float, int>> near;
    const float r2 = (float)((double)radius * (double)radius);
    for (int i = 0; i < n; ++i) {
        float dx = back.x - kp[i].x, dy = back.y - kp[i].y, dz = back.z - kp[i].z;
        float d = dx * dx; d += dy * dy; d += dz * dz;
        if (d < r2) near.emplace_back(d, i);
    }
    std::sort(near.begin(), near.end());
    // VoxelGrid(density) over the selected pose positions (same semantics as the cloud filter: ascending voxel index,
    // fp32 centroid), then nearest-1 over ALL key poses restores a real keyframe id (:991-997)
    if (!near.empty()) {
        const int m = (int)near.size();
        float mn[3] = {kp[near[0].second].x, kp[near[0].second].y, kp[near[0].second].z}, mx[3] = {mn[0], mn[1], mn[2]};
        for (auto& e : near) { const KeyPose& p = kp[e.second]; mn[0] = std::min(mn[0], p.x); mx[0] = std::max(mx[0], p.x); mn[1] = std::min(mn[1], p.y); mx[1] = std::max(mx[1], p.y); mn[2] = std::min(mn[2], p.z); mx[2] = std::max(mx[2], p.z); }
        const float inv = 1.0f / density;
        int min_b[3], div_b[3];
        for (int a = 0; a < 3; ++a) { min_b[a] = (int)std::floor(mn[a] * inv); div_b[a] = (int)std::floor(mx[a] * inv) - min_b[a] + 1; }
        int64_t dx = (int64_t)((mx[0] - mn[0]) * inv) + 1, dy = (int64_t)((mx[1] - mn[1]) * inv) + 1, dz = (int64_t)((mx[2] - mn[2]) * inv) + 1;
        std::vector<std::pair<int, int>> iv(m);
        const bool overflow = dx * dy * dz > (int64_t)INT32_MAX;
        for (int k = 0; k < m; ++k) {
            const KeyPose& p = kp[near[k].second];
            int i0 = (int)(std::floor(p.x * inv) - (float)min_b[0]), i1 = (int)(std::floor(p.y * inv) - (float)min_b[1]), i2 = (int)(std::floor(p.z * inv) - (float)min_b[2]);
            iv[k] = {overflow ? k : i0 + i1 * div_b[0] + i2 * div_b[0] * div_b[1], k};
        }
        std::sort(iv.begin(), iv.end());
        // candidates of the nearest-1 searches below (see there): the poses closer to `back` than shell_r, bucketed into cells of edge g = 2 density
        const double shell_r = (double)radius + 4.0 * (double)density + 1.0;
        const double g = 2.0 * (double)density;
        auto cell_of = [&](double v, double o) { return (long long)std::floor((v - o) / g) + (1ll << 19); };
        auto key_of = [](long long cx, long long cy, long long cz) { return (cz << 42) | (cy << 21) | cx; };
        std::vector<std::pair<long long, int>> cells;                                  // (cell key, pose index), sorted
        const bool grid_ok = g > 0.0 && shell_r < g * (double)(1 << 18);               // cell coordinates fit their 21 bits (else: full scans below)
        for (int i = 0; grid_ok && i < n; ++i) {
            const double ax = (double)kp[i].x - back.x, ay = (double)kp[i].y - back.y, az = (double)kp[i].z - back.z;
            const double q = ax * ax + ay * ay + az * az;
            if (q < shell_r * shell_r) cells.emplace_back(key_of(cell_of(kp[i].x, back.x), cell_of(kp[i].y, back.y), cell_of(kp[i].z, back.z)), i);   // (a NaN pose is left out: it can never win a `<`)
        }
        const bool shell_is_all = (int)cells.size() == n;
        std::sort(cells.begin(), cells.end());
        int k = 0;
        while (k < m) {
            int j = k; float sx = 0, sy = 0, sz = 0;
            while (j < m && iv[j].first == iv[k].first) { const KeyPose& p = kp[near[iv[j].second].second]; sx += p.x; sy += p.y; sz += p.z; ++j; }
            float c = (float)(j - k), cx = sx / c, cy = sy / c, cz = sz / c;
            // nearest-1 over ALL key poses, ties to the lower index (the order-independent form of "ascending scan, strict <").
            //  (1) the voxel's own members are key poses: the best of them bounds the answer, bd <= (voxel diagonal)^2;
            //  (2) any pose that beats or ties bd differs from the centroid by at most b = sqrt(bd) per axis, so it sits in one of the <= 2 x 2 x 2 cells around the
            //      centroid (cell edge g >= b; the same monotonic cell function is applied to both, b carries a 1e-4 margin against the fp32 rounding of d);
            //  (3) a pose OUTSIDE the shell is at least shell_r - |c - back| from the centroid (triangle inequality, in double, with margin): once bd is below that,
            //      nothing outside can win or tie.  Otherwise — never seen on a drive; publishGlobalMap's 1 km radius makes the shell everything anyway — every
            //      pose is scanned, as the reference's kd-tree would.
            int best = 0; float bd = INFINITY;
            auto consider = [&](int i) {
                float ex = cx - kp[i].x, ey = cy - kp[i].y, ez = cz - kp[i].z;
                float d = ex * ex; d += ey * ey; d += ez * ez;
                if (d < bd || (d == bd && i < best)) { bd = d; best = i; }
            };
            for (int t = k; t < j; ++t) consider(near[iv[t].second].second);
            bool done = false;
            if (grid_ok && bd < INFINITY && cx == cx && cy == cy && cz == cz) {
                const double b = std::sqrt((double)bd) * (1.0 + 1e-4) + 1e-30;
                if (b < g) {
                    const long long x0 = cell_of(cx - b, back.x), x1 = cell_of(cx + b, back.x), y0 = cell_of(cy - b, back.y), y1 = cell_of(cy + b, back.y),
                                    z0 = cell_of(cz - b, back.z), z1 = cell_of(cz + b, back.z);
                    for (long long zc = z0; zc <= z1; ++zc)
                        for (long long yc = y0; yc <= y1; ++yc) {
                            auto it = std::lower_bound(cells.begin(), cells.end(), std::make_pair(key_of(x0, yc, zc), -1));
                            const long long last = key_of(x1, yc, zc);
                            for (; it != cells.end() && it->first <= last; ++it) consider(it->second);
                        }
                    const double c_back = std::sqrt((double)(cx - back.x) * (cx - back.x) + (double)(cy - back.y) * (cy - back.y) + (double)(cz - back.z) * (cz - back.z));
                    const double outside = shell_r - c_back;                           // lower bound of |p - c| for every pose p outside the shell
                    done = shell_is_all || (outside > 0.0 && outside * outside * (1.0 - 1e-4) > (double)bd);
                }
            }
            if (!done) {
                best = 0; bd = INFINITY;
                for (int i = 0; i < n; ++i) consider(i);
            }
            if (!(point_distance(cx, cy, cz, back.x, back.y, back.z) > radius)) out.push_back(best);      // :1018 on the centroid
            k = j;
        }
    }
    if (include_recent)
        for (int i = n - 1; i >= 0; --i) {                                        // :1000-1007
            if (time_cur - kp[i].time < 10.0) { if (!(point_distance(kp[i].x, kp[i].y, kp[i].z, back.x, back.y, back.z) > radius)) out.push_back(i); }   // :1018
            else break;
        }
    return out;
}

}  // namespace liorf_host

// host_logic.hpp — the small scalar host-side steps around the GPU hot path, written as plain C++ (no CUDA, no ROS,
// no PCL).  These mirror reference code that is host logic in liorf too:
//   extract_nearby            mapOptimization::extractNearby            src/mapOptmization.cpp:975-1010 (keyframe SELECTION only)
//   save_frame                mapOptimization::saveFrame                src/mapOptmization.cpp:1365-1384
//   transform_update_clamp    mapOptimization::transformUpdate          src/mapOptmization.cpp:1348-1350 (the clamps)
// They run on a few hundred key poses per frame; the heavy half of extractSurroundingKeyFrames (extractCloud) is CUDA.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <utility>
#include <vector>

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
inline void transform_update_clamp(float tf[6], float rotation_tollerance, float z_tollerance) {                // :1348-1350
    tf[0] = constraint_transformation(tf[0], rotation_tollerance);
    tf[1] = constraint_transformation(tf[1], rotation_tollerance);
    tf[5] = constraint_transformation(tf[5], z_tollerance);
}

// extractNearby (:975-1010): ids of the keyframes whose clouds extractCloud will fuse, IN ORDER, duplicates included.
inline std::vector<int> extract_nearby(const std::vector<KeyPose>& kp, double time_cur, float radius, float density) {
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
        int k = 0;
        while (k < m) {
            int j = k; float sx = 0, sy = 0, sz = 0;
            while (j < m && iv[j].first == iv[k].first) { const KeyPose& p = kp[near[iv[j].second].second]; sx += p.x; sy += p.y; sz += p.z; ++j; }
            float c = (float)(j - k), cx = sx / c, cy = sy / c, cz = sz / c;
            int best = 0; float bd = INFINITY;
            for (int i = 0; i < n; ++i) {
                float ex = cx - kp[i].x, ey = cy - kp[i].y, ez = cz - kp[i].z;
                float d = ex * ex; d += ey * ey; d += ez * ez;
                if (d < bd) { bd = d; best = i; }
            }
            out.push_back(best);
            k = j;
        }
    }
    for (int i = n - 1; i >= 0; --i) {                                            // :1000-1007
        if (time_cur - kp[i].time < 10.0) out.push_back(i); else break;
    }
    return out;
}

}  // namespace liorf_host

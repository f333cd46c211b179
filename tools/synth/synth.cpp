// synth.cpp — deterministic synthetic inputs for the tests and the benchmark.
// TEST/BENCH INFRASTRUCTURE, not product code.  Everything is a pure function of (seed, arguments), so the CPU
// oracle and the GPU library always read identical bytes.
//
// Scene: an endless street lattice (pitch 40 m).  Every lattice cell holds one box building [6,34]x[6,34] m of hashed
// height 4..15 m (absent with probability 0.2 → gaps), and three vertical cylinders (poles/trunks, r 0.15..0.5 m, 6 m
// high) at hashed positions in the street margin.  Ground plane at z = -1.73 m (HDL-64 mounting height).
// Rays are traced by a 2-D DDA over lattice cells up to 120 m.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct PRaw { float x, y, z, i; uint16_t ring; uint16_t pad; float time; };

inline uint64_t mix(uint64_t x) { x += 0x9e3779b97f4a7c15ull; x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull; x = (x ^ (x >> 27)) * 0x94d049bb133111ebull; return x ^ (x >> 31); }
inline double u01(uint64_t h) { return (double)(mix(h) >> 11) * (1.0 / 9007199254740992.0); }
inline uint64_t cell_hash(int cx, int cy, int k) { return mix(((uint64_t)(uint32_t)cx << 32) ^ (uint64_t)(uint32_t)cy) ^ mix(0x1234567ull + (uint64_t)k * 7919ull); }

constexpr double PITCH = 40.0, GROUND_Z = -1.73, MAX_RANGE = 120.0, MIN_RANGE = 1.0;

// ray o + t d (d unit).  Returns nearest hit t in (0, MAX_RANGE] or -1.
double trace(const double o[3], const double d[3]) {
    double best = 1e30;
    if (d[2] < -1e-9) { double t = (GROUND_Z - o[2]) / d[2]; if (t > 0) best = t; }
    // 2-D DDA over lattice cells
    int cx = (int)std::floor(o[0] / PITCH), cy = (int)std::floor(o[1] / PITCH);
    const int sx = d[0] > 0 ? 1 : -1, sy = d[1] > 0 ? 1 : -1;
    const double ix = std::fabs(d[0]) > 1e-12 ? 1.0 / d[0] : 1e30, iy = std::fabs(d[1]) > 1e-12 ? 1.0 / d[1] : 1e30;
    double tmx = std::fabs(d[0]) > 1e-12 ? (((cx + (sx > 0)) * PITCH) - o[0]) * ix : 1e30;
    double tmy = std::fabs(d[1]) > 1e-12 ? (((cy + (sy > 0)) * PITCH) - o[1]) * iy : 1e30;
    double t_enter = 0.0;
    for (int step = 0; step < 16 && t_enter < std::min(best, MAX_RANGE); ++step) {
        const double bx = cx * PITCH, by = cy * PITCH;
        // building
        if (u01(cell_hash(cx, cy, 0)) < 0.8) {
            const double h = 4.0 + 11.0 * u01(cell_hash(cx, cy, 1));
            const double lo[3] = {bx + 6.0, by + 6.0, GROUND_Z}, hi[3] = {bx + 34.0, by + 34.0, GROUND_Z + h};
            double t0 = 0.0, t1 = 1e30; bool ok = true;
            for (int a = 0; a < 3 && ok; ++a) {
                if (std::fabs(d[a]) < 1e-12) { if (o[a] < lo[a] || o[a] > hi[a]) ok = false; }
                else { double ta = (lo[a] - o[a]) / d[a], tb = (hi[a] - o[a]) / d[a]; if (ta > tb) std::swap(ta, tb); t0 = std::max(t0, ta); t1 = std::min(t1, tb); if (t0 > t1) ok = false; }
            }
            if (ok && t0 > 1e-6 && t0 < best) best = t0;
        }
        // cylinders in the street margin
        for (int k = 0; k < 3; ++k) {
            const double r = 0.15 + 0.35 * u01(cell_hash(cx, cy, 10 + k));
            const double a = u01(cell_hash(cx, cy, 20 + k)), b = u01(cell_hash(cx, cy, 30 + k));
            // positions along the left / bottom street margins of the cell (2.5 m off the building line)
            const double px = (k == 0) ? bx + 3.5 : bx + 6.0 + 28.0 * a;
            const double py = (k == 0) ? by + 6.0 + 28.0 * b : (k == 1 ? by + 3.5 : by + 36.5);
            const double ox = o[0] - px, oy = o[1] - py;
            const double A = d[0] * d[0] + d[1] * d[1];
            if (A < 1e-12) continue;
            const double B = ox * d[0] + oy * d[1], Cc = ox * ox + oy * oy - r * r;
            const double disc = B * B - A * Cc;
            if (disc < 0) continue;
            const double t = (-B - std::sqrt(disc)) / A;
            if (t > 1e-6 && t < best) { const double z = o[2] + t * d[2]; if (z >= GROUND_Z && z <= GROUND_Z + 6.0) best = t; }
        }
        if (tmx < tmy) { t_enter = tmx; tmx += PITCH * std::fabs(ix); cx += sx; }
        else { t_enter = tmy; tmy += PITCH * std::fabs(iy); cy += sy; }
    }
    return (best <= MAX_RANGE && best >= MIN_RANGE) ? best : -1.0;
}

inline void rot_zyx(double roll, double pitch, double yaw, double R[9]) {       // R = Rz(yaw) Ry(pitch) Rx(roll)
    const double A = std::cos(yaw), B = std::sin(yaw), C = std::cos(pitch), D = std::sin(pitch), E = std::cos(roll), F = std::sin(roll);
    R[0] = A * C; R[1] = A * D * F - B * E; R[2] = B * F + A * D * E;
    R[3] = B * C; R[4] = A * E + B * D * F; R[5] = B * D * E - A * F;
    R[6] = -D;    R[7] = C * F;             R[8] = C * E;
}

}  // namespace

extern "C" {

// sensor: 0 = HDL-64 (64 x 1875, -24.8..+2 deg), 1 = OS1-128 (128 x 1024, +-22.5 deg), 2 = Livox rosette (24000 pts)
// pose6 = (roll,pitch,yaw,x,y,z) of the sensor in the world at scan start; omega = body angular rate (rad/s) applied
// during the 0.1 s sweep (motion distortion for the deskew tests); vel = world-frame linear velocity (m/s).
// Points are expressed in the sensor frame AT THEIR OWN capture time (i.e. distorted).  Returns the number of returns.
int synth_scan(int sensor, const double* pose6, const double* omega, const double* vel, uint64_t seed, double noise_sigma, PRaw* out, int capacity) {
    int rings, cols; double el0, el1;
    if (sensor == 0) { rings = 64; cols = 1875; el0 = -24.8; el1 = 2.0; }
    else if (sensor == 1) { rings = 128; cols = 1024; el0 = -22.5; el1 = 22.5; }
    else { rings = 6; cols = 4000; el0 = -12.55; el1 = 12.55; }
    const int total = rings * cols;
    std::vector<PRaw> tmp(total); std::vector<uint8_t> ok(total);
    const double sweep = 0.1;
#pragma omp parallel for schedule(static, 2048)
    for (int i = 0; i < total; ++i) {
        int ring, col; double az, el, t;
        if (sensor == 2) {
            ring = i % 6; col = i / 6; t = sweep * (double)i / total;
            const double tau = (double)i / total;
            az = 40.85 * std::sin(2 * M_PI * 17.0 * tau) * M_PI / 180.0;
            el = 12.55 * std::sin(2 * M_PI * 17.0 * 1.6180339887 * tau + 0.3 * ring) * M_PI / 180.0;
        } else {
            col = i / rings; ring = i % rings; t = sweep * (double)col / cols;
            az = -2 * M_PI * (double)col / cols;                                  // clockwise sweep like a spinning lidar
            el = (el0 + (el1 - el0) * (double)ring / (rings - 1)) * M_PI / 180.0;
        }
        // sensor pose at time t
        double R[9]; rot_zyx(pose6[0] + omega[0] * t, pose6[1] + omega[1] * t, pose6[2] + omega[2] * t, R);
        const double o[3] = {pose6[3] + vel[0] * t, pose6[4] + vel[1] * t, pose6[5] + vel[2] * t};
        const double ds[3] = {std::cos(el) * std::cos(az), std::cos(el) * std::sin(az), std::sin(el)};
        const double dw[3] = {R[0] * ds[0] + R[1] * ds[1] + R[2] * ds[2], R[3] * ds[0] + R[4] * ds[1] + R[5] * ds[2], R[6] * ds[0] + R[7] * ds[1] + R[8] * ds[2]};
        double r = trace(o, dw);
        ok[i] = 0;
        if (r > 0) {
            // Box-Muller range noise
            const double u1 = std::max(u01(seed * 0x100000001b3ull + (uint64_t)i * 2 + 1), 1e-12), u2 = u01(seed * 0x100000001b3ull + (uint64_t)i * 2 + 2);
            r += noise_sigma * std::sqrt(-2.0 * std::log(u1)) * std::cos(2 * M_PI * u2);
            if (r >= MIN_RANGE && r <= MAX_RANGE) {
                PRaw p; p.x = (float)(r * ds[0]); p.y = (float)(r * ds[1]); p.z = (float)(r * ds[2]);
                p.i = (float)(10.0 + 90.0 * u01(seed + 77 + (uint64_t)i)); p.ring = (uint16_t)ring; p.pad = 0; p.time = (float)t;
                tmp[i] = p; ok[i] = 1;
            }
        }
    }
    int n = 0;
    for (int i = 0; i < total && n < capacity; ++i) if (ok[i]) out[n++] = tmp[i];
    return n;
}

int synth_scan_capacity(int sensor) { return sensor == 0 ? 64 * 1875 : (sensor == 1 ? 128 * 1024 : 24000); }

// ScanContext database of config 5: per entry a smooth random height field over (ring, sector): sum of 6 random
// low-frequency 2-D sinusoids scaled to [0, 12] m, 25 % empty cells (0.0).  desc: count x 1200 doubles, row-major.
void synth_sc_descriptors(uint64_t seed, int first, int count, double* desc) {
#pragma omp parallel for schedule(static, 64)
    for (int e = 0; e < count; ++e) {
        const uint64_t id = seed * 1315423911ull + (uint64_t)(first + e);
        double a[6], fr[6], fs[6], ph[6];
        for (int k = 0; k < 6; ++k) {
            a[k] = 0.5 + u01(id * 64 + k); fr[k] = 0.2 + 1.3 * u01(id * 64 + 8 + k); fs[k] = (double)(1 + (int)(4 * u01(id * 64 + 16 + k)));
            ph[k] = 2 * M_PI * u01(id * 64 + 24 + k);
        }
        double* d = desc + (size_t)e * 1200;
        for (int r = 0; r < 20; ++r) for (int s = 0; s < 60; ++s) {
            double v = 0, amp = 0;
            for (int k = 0; k < 6; ++k) { v += a[k] * std::sin(fr[k] * r * 0.6 + fs[k] * s * (2 * M_PI / 60.0) + ph[k]); amp += a[k]; }
            v = 6.0 + 6.0 * v / amp;
            const bool empty = u01(id * 4096 + 100 + r * 60 + s) < 0.25;
            d[r * 60 + s] = empty ? 0.0 : v;
        }
    }
}

// Query set of config 5: even queries = database entry src[q] column-shifted by shift[q] plus N(0, 0.05) on non-empty
// cells (true loops); odd queries = fresh descriptors.  db: K x 1200.
void synth_sc_queries(uint64_t seed, const double* db, int K, int Q, double* qdesc, int* src, int* shift) {
    for (int q = 0; q < Q; ++q) {
        double* d = qdesc + (size_t)q * 1200;
        if (q % 2 == 0) {
            const int e = (int)(u01(seed * 31 + 5 + (uint64_t)q * 3) * K) % K, s = (int)(u01(seed * 31 + 6 + (uint64_t)q * 3) * 60) % 60;
            src[q] = e; shift[q] = s;
            const double* b = db + (size_t)e * 1200;
            for (int r = 0; r < 20; ++r) for (int c = 0; c < 60; ++c) {
                double v = b[r * 60 + c];
                if (v != 0.0) {
                    const double u1 = std::max(u01(seed + 999 + (uint64_t)q * 2400 + r * 120 + c * 2), 1e-12), u2 = u01(seed + 999 + (uint64_t)q * 2400 + r * 120 + c * 2 + 1);
                    v += 0.05 * std::sqrt(-2.0 * std::log(u1)) * std::cos(2 * M_PI * u2);
                }
                d[r * 60 + (c + s) % 60] = v;
            }
        } else {
            src[q] = -1; shift[q] = 0;
            synth_sc_descriptors(seed ^ 0xabcdef12345ull, 1000000 + q, 1, d);
        }
    }
}

}  // extern "C"

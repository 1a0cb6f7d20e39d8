// dronechase_b200 -- projection "LiDAR": entity centres -> 13x26 spherical grid.
//
// The reference's sensor is not a ray cast: both classes project KNOWN entity positions into the
// observer's body frame and bin them (fused_lidar.py:143-150 docstring).  Bit-exact targets:
//   fused   FusedLIDAR.update_data fused_lidar.py:143-217 + lidar_math.py:25-34 (cartesian_to_spherical),
//           :53-83 (reframe), :94-96 (index_from_radian: truncation + clip), :128-129, :274-311 (add_features:
//           float64 challenger vs float32 cell, strict '<')
//   classic LIDAR._add_end_position/_add_spherical/_normalize_angle lidar.py:151-200,290-308 (cull unless
//           0<r<R, Python round() modulo n, '>' rejects so the last equal entity wins)
// Cell indices and winners are decided in float64 from the float32 snapshot, like the reference.
#pragma once
#include "common.cuh"

namespace dc {

struct LidarHit {
    int cell;      // theta_idx * 26 + phi_idx, or -1 when the entity cannot mark a cell
    double rn;     // normalised distance (float64, as the reference compares it)
    double theta, phi;   // continuous angles (FusedLIDAR keeps them in its feature list)
};

// own_p/own_q and p are the float32 snapshot values (perception_snapshot.py:91-110) for the fused
// flavour; the classic flavour consumes the float64 message directly, so pass the full precision.
__device__ __forceinline__ LidarHit lidar_project_one(int flavour, double radius,
                                                      double opx, double opy, double opz,
                                                      double oqx, double oqy, double oqz, double oqw,
                                                      double px, double py, double pz) {
    const double PI = 3.141592653589793;
    double ix, iy, iz, iw;      // inverse rotation quaternion
    if (flavour == 0) {
        // LidarMath._invert_quaternion in float32 arithmetic (the snapshot arrays are float32)
        float fx = (float)oqx, fy = (float)oqy, fz = (float)oqz, fw = (float)oqw;
        float nsq = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(fx, fx), __fmul_rn(fy, fy)), __fmul_rn(fz, fz)), __fmul_rn(fw, fw));
        ix = (double)__fdiv_rn(-fx, nsq); iy = (double)__fdiv_rn(-fy, nsq);
        iz = (double)__fdiv_rn(-fz, nsq); iw = (double)__fdiv_rn(fw, nsq);
    } else {
        ix = -oqx; iy = -oqy; iz = -oqz; iw = oqw;     // getMatrixFromQuaternion(q)^T
    }
    const double dx = px - opx, dy = py - opy, dz = pz - opz;
    const double r00 = 1 - 2 * (iy * iy + iz * iz), r01 = 2 * (ix * iy - iw * iz), r02 = 2 * (ix * iz + iw * iy);
    const double r10 = 2 * (ix * iy + iw * iz), r11 = 1 - 2 * (ix * ix + iz * iz), r12 = 2 * (iy * iz - iw * ix);
    const double r20 = 2 * (ix * iz - iw * iy), r21 = 2 * (iy * iz + iw * ix), r22 = 1 - 2 * (ix * ix + iy * iy);
    const double x = r00 * dx + r01 * dy + r02 * dz;
    const double y = r10 * dx + r11 * dy + r12 * dz;
    const double z = r20 * dx + r21 * dy + r22 * dz;
    const double r = sqrt(x * x + y * y + z * z);
    double theta = 0.0, phi = 0.0;
    if (r != 0.0) {
        theta = acos(fmin(fmax(z / r, -1.0), 1.0));
        phi = atan2(y, x);
    }
    LidarHit h;
    h.theta = theta; h.phi = phi;
    if (flavour == 0) {
        h.rn = fmin(fmax(r / radius, 0.0), 1.0);
        int ti = (int)(theta / PI * N_THETA);
        int pj = (int)((phi + PI) / (2 * PI) * N_PHI);
        ti = min(max(ti, 0), N_THETA - 1);
        pj = min(max(pj, 0), N_PHI - 1);
        h.cell = ti * N_PHI + pj;
    } else {
        if (!(r > 0.0 && r < radius)) { h.cell = -1; h.rn = 1.0; return h; }
        h.rn = r / radius;
        // Python round() is round-half-even == rint() in the default rounding mode
        int ti = ((int)rint(theta / PI * N_THETA)) % N_THETA;
        int pj = ((int)rint((phi + PI) / (2 * PI) * N_PHI)) % N_PHI;
        h.cell = ti * N_PHI + pj;
    }
    return h;
}

// float64 angles -> cell indices, exactly as lidar_project_one (taken for ~1e-3 of the entities)
__device__ __forceinline__ void lidar_cell_slow(double x, double y, double z, double r, int* ti, int* pj) {
    const double PI = 3.141592653589793;
    double theta = 0.0, phi = 0.0;
    if (r != 0.0) {
        theta = acos(fmin(fmax(z / r, -1.0), 1.0));
        phi = atan2(y, x);
    }
    *ti = (int)(theta / PI * N_THETA);
    *pj = (int)((phi + PI) / (2 * PI) * N_PHI);
}

// Fused flavour when only (cell, r_n) are wanted (level4: the features are not kept).  The float64 acos/atan2 of
// lidar_project_one only decide a cell index, so the angles are first taken in float32 (good to ~1e-6 rad = 4e-6
// cells; acos is well conditioned at every interior border k pi / 13) and the float64 path runs only when one of them
// lies within 2e-3 of a cell border -- the cell is the one the reference's float64 arithmetic picks either way.
// r_n is float64 as in the reference.
__device__ __forceinline__ void lidar_cell_fused(double radius, double opx, double opy, double opz,
                                                 double oqx, double oqy, double oqz, double oqw,
                                                 double px, double py, double pz, int* cell, double* rn_out) {
    float fx = (float)oqx, fy = (float)oqy, fz = (float)oqz, fw = (float)oqw;
    float nsq = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(fx, fx), __fmul_rn(fy, fy)), __fmul_rn(fz, fz)), __fmul_rn(fw, fw));
    const double ix = (double)__fdiv_rn(-fx, nsq), iy = (double)__fdiv_rn(-fy, nsq);
    const double iz = (double)__fdiv_rn(-fz, nsq), iw = (double)__fdiv_rn(fw, nsq);
    const double dx = px - opx, dy = py - opy, dz = pz - opz;
    const double r00 = 1 - 2 * (iy * iy + iz * iz), r01 = 2 * (ix * iy - iw * iz), r02 = 2 * (ix * iz + iw * iy);
    const double r10 = 2 * (ix * iy + iw * iz), r11 = 1 - 2 * (ix * ix + iz * iz), r12 = 2 * (iy * iz - iw * ix);
    const double r20 = 2 * (ix * iz - iw * iy), r21 = 2 * (iy * iz + iw * ix), r22 = 1 - 2 * (ix * ix + iy * iy);
    const double x = r00 * dx + r01 * dy + r02 * dz;
    const double y = r10 * dx + r11 * dy + r12 * dz;
    const double z = r20 * dx + r21 * dy + r22 * dz;
    const double r = sqrt(x * x + y * y + z * z);
    *rn_out = fmin(fmax(r / radius, 0.0), 1.0);
    int ti = 0, pj = 0;
    bool fast = false;
    if (r > 1e-9) {
        const float k = 4.1380285203892786f;              // 13 / pi = 26 / (2 pi)
        const float tf = acosf(fminf(fmaxf(__fdividef((float)z, (float)r), -1.0f), 1.0f)) * k;
        const float pf = (atan2f((float)y, (float)x) + 3.14159265358979f) * k;
        const float ft = tf - floorf(tf), fp = pf - floorf(pf), m = 2e-3f;
        fast = ft > m && ft < 1.0f - m && fp > m && fp < 1.0f - m;
        ti = (int)tf; pj = (int)pf;
    }
    if (!fast) lidar_cell_slow(x, y, z, r, &ti, &pj);
    ti = min(max(ti, 0), N_THETA - 1);
    pj = min(max(pj, 0), N_PHI - 1);
    *cell = ti * N_PHI + pj;
}

// The float32 product build of env_kernel: the whole projection in float32 (about 120 four-cycle instructions instead of a
// float64 chain with a square root, a division and two libm calls that cost a warp ~3.4 us per trip), the exact float64
// path only when an angle lies within 2e-4 of a cell border.  The float32 angles are good to ~4e-6 cells (inputs are
// float32 snapshot values; acos is well conditioned at every interior border k pi / 13, the two poles are clipped), so
// the cell is the one the reference's float64 arithmetic picks; r_n carries float32 rounding (<= 2 ulp of the value the
// reference stores as float32).
__device__ __noinline__ int lidar_cell_exact(double radius, double opx, double opy, double opz,
                                             double oqx, double oqy, double oqz, double oqw, double px, double py, double pz) {
    int c; double rn;
    lidar_cell_fused(radius, opx, opy, opz, oqx, oqy, oqz, oqw, px, py, pz, &c, &rn);
    return c;
}
__device__ __forceinline__ void lidar_cell_fused_f32(float radius, float opx, float opy, float opz, float fx, float fy, float fz, float fw,
                                                     float px, float py, float pz, int* cell, float* rn_out) {
    const float inv = mufu_rcp(fmaf(fw, fw, fmaf(fz, fz, fmaf(fy, fy, fx * fx))));      // LidarMath._invert_quaternion
    const float ix = -fx * inv, iy = -fy * inv, iz = -fz * inv, iw = fw * inv;
    const float dx = px - opx, dy = py - opy, dz = pz - opz;
    const float x2 = ix + ix, y2 = iy + iy, z2 = iz + iz;
    const float xx = x2 * ix, yy = y2 * iy, zz = z2 * iz, xy = x2 * iy, xz = x2 * iz, yz = y2 * iz;
    const float x = fmaf(fmaf(y2, iw, xz), dz, fmaf(fmaf(-z2, iw, xy), dy, (1.0f - yy - zz) * dx));
    const float y = fmaf(fmaf(-x2, iw, yz), dz, fmaf(1.0f - xx - zz, dy, fmaf(z2, iw, xy) * dx));
    const float z = fmaf(1.0f - xx - yy, dz, fmaf(fmaf(x2, iw, yz), dy, fmaf(-y2, iw, xz) * dx));
    const float r2 = fmaf(z, z, fmaf(y, y, x * x));
    int ti = 0, pj = 0;
    bool fast = false;
    float r = 0.0f;
    if (r2 > 1e-18f) {
        const float ir = mufu_rsq(r2);
        r = r2 * ir;
        const float k = 4.1380285203892786f;              // 13 / pi = 26 / (2 pi)
        const float tf = acosf(fminf(fmaxf(z * ir, -1.0f), 1.0f)) * k;
        const float pf = (atan2f(y, x) + 3.14159265358979f) * k;
        const float ft = tf - floorf(tf), fp = pf - floorf(pf), m = 2e-4f;
        fast = ft > m && ft < 1.0f - m && fp > m && fp < 1.0f - m;
        ti = (int)tf; pj = (int)pf;
    }
    *rn_out = fminf(r * mufu_rcp(radius), 1.0f);
    if (fast) {
        ti = min(max(ti, 0), N_THETA - 1);
        pj = min(max(pj, 0), N_PHI - 1);
        *cell = ti * N_PHI + pj;
    } else {
        *cell = lidar_cell_exact((double)radius, (double)opx, (double)opy, (double)opz, (double)fx, (double)fy, (double)fz, (double)fw,
                                 (double)px, (double)py, (double)pz);
    }
}

// Sequential add_features/_add_spherical over the entity list, evaluated from entity k's point of
// view: returns true when k is the entity whose values the cell finally holds.
__device__ __forceinline__ bool lidar_wins(int flavour, int k, int n, const int* cells, const double* rns) {
    const int c = cells[k];
    if (c < 0) return false;
    float cur = 1.0f;
    int win = -1;
    for (int j = 0; j < n; ++j) {
        if (cells[j] != c) continue;
        const double rn = rns[j];
        const bool take = (flavour == 0) ? (rn < (double)cur) : !(rn > (double)cur);
        if (take) { cur = (float)rn; win = j; }
    }
    return win == k;
}

}  // namespace dc

// dronechase_b200 -- QuadX drone dynamics, one thread per ARMED drone, state in registers.
//
// Replaces, per physics substep and per armed drone, the call sequence of
//   level4_simulation.py:84-98   update_imu -> update_control -> update_physics ; stepSimulation
// i.e. PyFlyt QuadX.update_state / update_control (mode 6 cascade) / update_physics (Motors,
// BoringBodies) and Bullet's integration of one free rigid body (pyflyt==0.11.1, pybullet==3.2.7,
// neither vendored in the reference; the arithmetic restated here is documented in DESIGN.md and
// mirrored by the float64 oracle oracle/dynamics.py).
#pragma once
#include "common.cuh"

namespace dc {

enum { PID_ANG_VEL = 0, PID_ANG_POS = 1, PID_LIN_VEL = 2, PID_Z_VEL = 3, PID_LIN_POS = 4, PID_Z_POS = 5 };

template <typename R> struct QuadParams {
    R mass, inv_mass, inertia[3], inv_inertia[3], arm, kf, km, dt_over_tau, noise_ratio, max_rpm;
    R drag_k, dt, pid_T, inv_pid_T, gravity, ground_z;
    int gyro;
    R kp[6][3], ki[6][3], kd[6][3], lim[6][3];
    R kiT[6][3], kdiT[6][3];   // ki * pid_T and kd / pid_T, folded on the host
};

template <typename R> struct Drone {
    R px, py, pz;          // world position
    R qx, qy, qz, qw;      // body->world quaternion (PyBullet order x,y,z,w)
    R vx, vy, vz;          // world linear velocity
    R wx, wy, wz;          // body angular velocity
    R thr[4];              // motor throttle
    R pid[24];             // PyFlyt PID integrators / previous errors (PID_SLOTS; words 18..23 belong to mode 7)
};

// What the reference's IMU publishes (imu.py:27-41): body-frame velocities, euler, world position.
template <typename R> struct Imu {
    R px, py, pz, roll, pitch, yaw, ub, vb, wb, p, q, r;
    R qx, qy, qz, qw;
};

// PyFlyt PID.step: integral += ki*err*T (clipped), derivative = kd*(err - prev)/T, output clipped.  ki*T and kd/T are
// folded on the host, so one controller is 2 adds, 3 FMAs and 4 min/max.
template <typename R>
__device__ __forceinline__ R pid_step(R& integ, R& prev, R kp, R kiT, R kdiT, R lim, R state, R sp) {
    const R err = sp - state;
    integ = clamp_(fma_(kiT, err, integ), -lim, lim);
    const R diff = err - prev;
    prev = err;
    return clamp_(fma_(kdiT, diff, fma_(kp, err, integ)), -lim, lim);
}
#define DC_PID(i, j) P.kp[i][j], P.kiT[i][j], P.kdiT[i][j], P.lim[i][j]

// Box-Muller pieces.  The float path uses the SFU intrinsics: the sample only scales a 2 % throttle
// perturbation, so 1e-6 absolute error is far below the float32 state resolution.
__device__ __forceinline__ float bm_radius(float u) {
    const float x = -1.3862943611198906f * __log2f(u);     // -2 ln u, u in (0,1] -> x >= 0
    return x * rsqrtf(fmaxf(x, 1e-30f));
}
__device__ __forceinline__ double bm_radius(double u) { return sqrt(-2.0 * log(u)); }
// angle 2 pi u of the uniform u = k / 2^24 given as the integer k
__device__ __forceinline__ void bm_angle(uint32_t k, float* s, float* c) { __sincosf((float)k * (float)(6.283185307179586 / 16777216.0), s, c); }
__device__ __forceinline__ void bm_angle(uint32_t k, double* s, double* c) { sincospi((double)k * (2.0 / 16777216.0), s, c); }

// Four standard normals for (env, drone slot, physics substep): Box-Muller on one Philox block.
template <typename R>
__device__ __forceinline__ void motor_noise(uint32_t k0, uint32_t k1, uint32_t env, uint32_t slot,
                                            uint32_t phys_step, R n[4]) {
    uint4 x = philox4x32_10(phys_step, slot * 256u + (uint32_t)STREAM_MOTOR, env, 0u, k0, k1);
    const R inv24 = (R)(1.0 / 16777216.0);
    const R u1 = fma_((R)(x.x >> 8), inv24, inv24), u3 = fma_((R)(x.z >> 8), inv24, inv24);   // (k + 1) / 2^24, exact
    R r1 = bm_radius(u1), r2 = bm_radius(u3);
    R s, c;
    bm_angle(x.y >> 8, &s, &c); n[0] = r1 * c; n[1] = r1 * s;
    bm_angle(x.w >> 8, &s, &c); n[2] = r2 * c; n[3] = r2 * s;
}

// atan2(a, b) for the roll angle.  In controlled flight |a| << b (the velocity loop limits the angle
// command to 0.4 rad), where the odd series up to t^15 in t = a/b is exact to float32 rounding
// (remainder 0.45^17/17 < 8e-8); anything else takes the library path.
__device__ __forceinline__ float atan2_small(float a, float b) {
    if (b > 0.0f && fabsf(a) <= 0.45f * b) {
        const float t = __fdividef(a, b), t2 = t * t;
        float p = -1.0f / 15.0f;
        p = fmaf(p, t2, 1.0f / 13.0f); p = fmaf(p, t2, -1.0f / 11.0f); p = fmaf(p, t2, 1.0f / 9.0f);
        p = fmaf(p, t2, -1.0f / 7.0f); p = fmaf(p, t2, 1.0f / 5.0f); p = fmaf(p, t2, -1.0f / 3.0f);
        return fmaf(p * t2, t, t);
    }
    return atan2f(a, b);
}
__device__ __forceinline__ double atan2_small(double a, double b) { return atan2(a, b); }

// asin for the pitch angle: |x| <= 0.5 in controlled flight (angle commands are limited to 0.4 rad), where
// x + x^3 P(x^2) with a weighted least-squares quartic P is within 0.7 ulp of asin (fit: DESIGN.md section 3).
__device__ __forceinline__ float asin_small(float x) {
    if (fabsf(x) <= 0.5f) {
        const float t = x * x;
        float p = 0.042095039f;
        p = fmaf(p, t, 0.024220509f); p = fmaf(p, t, 0.045462500f); p = fmaf(p, t, 0.074953534f); p = fmaf(p, t, 0.16666752f);
        return fmaf(p * t, x, x);
    }
    return asinf(x);
}
__device__ __forceinline__ double asin_small(double x) { return asin(x); }

// body->world rotation matrix of a unit quaternion (x, y, z, w)
template <typename R> struct Rot { R r00, r01, r02, r10, r11, r12, r20, r21, r22; };
template <typename R> __device__ __forceinline__ Rot<R> quat_rot(R x, R y, R z, R w) {
    const R xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
    return Rot<R>{1 - 2 * (yy + zz), 2 * (xy - wz), 2 * (xz + wy),
                  2 * (xy + wz), 1 - 2 * (xx + zz), 2 * (yz - wx),
                  2 * (xz - wy), 2 * (yz + wx), 1 - 2 * (xx + yy)};
}

// yaw of btQuaternion::getEulerZYX, needed as an angle only for the observation
template <typename R> __device__ __forceinline__ R quat_yaw(R x, R y, R z, R w) {
    const R sarg = (R)-2 * (x * z - w * y);
    if (sarg <= (R)-0.99999) return 2 * atan2_(x, -y);
    if (sarg >= (R)0.99999) return 2 * atan2_(-x, y);
    return atan2_(2 * (x * y + w * z), w * w + x * x - y * y - z * z);
}

// One physics substep.  sp = mode-6 setpoint (vx, vy, yaw-rate, vz) in the ground frame
// (quadcopter.py:408-413).  Fills `imu` with the state the reference's IMU reads at the START of
// the substep (update_imu precedes control and integration); imu.yaw is left to the caller
// (quat_yaw of imu.q*), the control law only needs sin/cos of it.
// Body-frame force/torque of one update_control + update_physics, and the rotation they act through.
template <typename R> struct Wrench { R fx, fy, fz, tx, ty, tz; };

// update_imu -> update_control -> update_physics of one drone: advances the PID and motor state and
// returns the body-frame wrench Bullet will integrate at the next stepSimulation.  mode7 = PyFlyt
// position mode (x, y, yaw, z setpoint: lin_pos / z_pos / yaw loops in front of the mode-6 cascade),
// used by the stage01 munition (level2/components/quadcopter_manager.py:68).
template <typename R, bool NOISE>
__device__ __forceinline__ Wrench<R> quad_forces(Drone<R>& s, const R sp_in[4], bool mode7, const QuadParams<R>& P,
                                                 Imu<R>& imu, uint32_t k0, uint32_t k1, uint32_t env,
                                                 uint32_t slot, uint32_t phys_step, const Rot<R>& M) {
    // ---- QuadX.update_state ------------------------------------------------------------------
    const R x = s.qx, y = s.qy, z = s.qz, w = s.qw;
    const R ub = M.r00 * s.vx + M.r10 * s.vy + M.r20 * s.vz;     // R^T v
    const R vb = M.r01 * s.vx + M.r11 * s.vy + M.r21 * s.vz;
    const R wb = M.r02 * s.vx + M.r12 * s.vy + M.r22 * s.vz;
    // btQuaternion::getEulerZYX; cos/sin(yaw) come straight from the atan2 arguments
    const R sarg = -M.r20;
    R roll, pitch, sy, cy;
    if (sarg <= (R)-0.99999 || sarg >= (R)0.99999) {      // gimbal-lock branch of Bullet
        pitch = sarg < 0 ? (R)-1.5707963267948966 : (R)1.5707963267948966; roll = 0;
        const R yaw = sarg < 0 ? 2 * atan2_(x, -y) : 2 * atan2_(-x, y);
        sincos_(yaw, &sy, &cy);
    } else {
        pitch = asin_small(sarg);
        roll = atan2_small(M.r21, w * w - x * x - y * y + z * z);
        const R ys = M.r10, yc = w * w + x * x - y * y - z * z;
        const R h2 = ys * ys + yc * yc;
        const R ih = h2 > 0 ? rsqrt_(h2) : 0;
        sy = ys * ih; cy = h2 > 0 ? yc * ih : (R)1;
    }
    imu.px = s.px; imu.py = s.py; imu.pz = s.pz;
    imu.roll = roll; imu.pitch = pitch;
    imu.ub = ub; imu.vb = vb; imu.wb = wb;
    imu.p = s.wx; imu.q = s.wy; imu.r = s.wz;
    imu.qx = x; imu.qy = y; imu.qz = z; imu.qw = w;

    // ---- QuadX.update_control: mode 7 front end, then the mode 6 cascade ----------------------------
    R* pid = s.pid;
    R sp[4] = {sp_in[0], sp_in[1], sp_in[2], sp_in[3]};
    if (mode7) {
        sp[0] = pid_step(pid[18], pid[20], DC_PID(4, 0), s.px, sp_in[0]);
        sp[1] = pid_step(pid[19], pid[21], DC_PID(4, 1), s.py, sp_in[1]);
        sp[3] = pid_step(pid[22], pid[23], DC_PID(5, 0), s.pz, sp_in[3]);
        const R yaw = quat_yaw(x, y, z, w);
        sp[2] = pid_step(pid[8], pid[11], DC_PID(1, 2), yaw, sp_in[2]);
    }
    const R u_cmd = cy * sp[0] + sy * sp[1];
    const R v_cmd = -sy * sp[0] + cy * sp[1];
    const R o0 = pid_step(pid[12], pid[14], DC_PID(2, 0), ub, u_cmd);
    const R o1 = pid_step(pid[13], pid[15], DC_PID(2, 1), vb, v_cmd);
    const R roll_cmd = -o1, pitch_cmd = o0;
    const R p_cmd = pid_step(pid[6], pid[9], DC_PID(1, 0), roll, roll_cmd);
    const R q_cmd = pid_step(pid[7], pid[10], DC_PID(1, 1), pitch, pitch_cmd);
    const R r_cmd = sp[2];
    const R tx = pid_step(pid[0], pid[3], DC_PID(0, 0), s.wx, p_cmd);
    const R ty = pid_step(pid[1], pid[4], DC_PID(0, 1), s.wy, q_cmd);
    const R tz = pid_step(pid[2], pid[5], DC_PID(0, 2), s.wz, r_cmd);
    R th = pid_step(pid[16], pid[17], DC_PID(3, 0), wb, sp[3]);
    th = clamp_(th, (R)0, (R)1);
    // motor mixing + saturation handling
    R pwm[4] = {-tx - ty + tz + th, tx + ty + tz + th, -tx + ty - tz + th, tx - ty - tz + th};
    const R high = max_(max_(pwm[0], pwm[1]), max_(pwm[2], pwm[3]));
    if constexpr (sizeof(R) == 4) {
        // branch-free: an unsaturated mix (the usual case) is multiplied by exactly 1 and moved by exactly 0
        const R ih = high > (R)1 ? fast_rcp(high) : (R)1;
#pragma unroll
        for (int m = 0; m < 4; ++m) pwm[m] *= ih;
        const R low = min_(min_(pwm[0], pwm[1]), min_(pwm[2], pwm[3]));
        const R kl = low < (R)0.05 ? ((R)0.05 - low) * fast_rcp((R)1 - low) : (R)0;
#pragma unroll
        for (int m = 0; m < 4; ++m) pwm[m] = fma_((R)1 - pwm[m], kl, pwm[m]);
    } else {
        if (high > (R)1) {
#pragma unroll
            for (int m = 0; m < 4; ++m) pwm[m] = pwm[m] / high;
        }
        const R low = min_(min_(pwm[0], pwm[1]), min_(pwm[2], pwm[3]));
        if (low < (R)0.05) {
#pragma unroll
            for (int m = 0; m < 4; ++m) pwm[m] = pwm[m] + ((R)1 - pwm[m]) / ((R)1 - low) * ((R)0.05 - low);
        }
    }

    // ---- QuadX.update_physics: Motors + BoringBodies ---------------------------------------------
    R nz[4] = {0, 0, 0, 0};
    if (NOISE) motor_noise<R>(k0, k1, env, slot, phys_step, nz);
    R thrust[4], fz = 0, mz = 0;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        R t = s.thr[m] + P.dt_over_tau * (pwm[m] - s.thr[m]);
        if (NOISE) t = t + nz[m] * t * P.noise_ratio;
        s.thr[m] = t;
        const R rpm = t * P.max_rpm, rpm2 = rpm * rpm;
        thrust[m] = P.kf * rpm2;
        fz += thrust[m];
        const R react = P.km * rpm2;
        mz += (m < 2) ? react : -react;
    }
    // propeller x/y signs consistent with the motor map: m0 (+,-) m1 (-,+) m2 (-,-) m3 (+,+)
    R tau_x = P.arm * (-thrust[0] + thrust[1] - thrust[2] + thrust[3]);
    R tau_y = -P.arm * (thrust[0] - thrust[1] - thrust[2] + thrust[3]);
    R tau_z = mz;
    const R fbx = -copysign(P.drag_k * ub * ub, ub);
    const R fby = -copysign(P.drag_k * vb * vb, vb);
    const R fbz = -copysign(P.drag_k * wb * wb, wb) + fz;
    return Wrench<R>{fbx, fby, fbz, tau_x, tau_y, tau_z};
}

// stepSimulation for one free rigid body: semi-implicit Euler, dt = 1/240, static plane clamp.
template <typename R>
__device__ __forceinline__ void quad_integrate(Drone<R>& s, const Wrench<R>& W, const QuadParams<R>& P, const Rot<R>& M) {
    const R x = s.qx, y = s.qy, z = s.qz, w = s.qw;
    const R r00 = M.r00, r01 = M.r01, r02 = M.r02, r10 = M.r10, r11 = M.r11, r12 = M.r12, r20 = M.r20, r21 = M.r21, r22 = M.r22;
    const R fbx = W.fx, fby = W.fy, fbz = W.fz;
    R tau_x = W.tx, tau_y = W.ty, tau_z = W.tz;
    const R dt = P.dt;
    const R ax = (r00 * fbx + r01 * fby + r02 * fbz) * P.inv_mass;
    const R ay = (r10 * fbx + r11 * fby + r12 * fbz) * P.inv_mass;
    const R az = (r20 * fbx + r21 * fby + r22 * fbz) * P.inv_mass + P.gravity;
    s.vx += dt * ax; s.vy += dt * ay; s.vz += dt * az;
    if (P.gyro) {
        const R Ix = P.inertia[0] * s.wx, Iy = P.inertia[1] * s.wy, Iz = P.inertia[2] * s.wz;
        tau_x -= s.wy * Iz - s.wz * Iy;
        tau_y -= s.wz * Ix - s.wx * Iz;
        tau_z -= s.wx * Iy - s.wy * Ix;
    }
    s.wx += dt * tau_x * P.inv_inertia[0];
    s.wy += dt * tau_y * P.inv_inertia[1];
    s.wz += dt * tau_z * P.inv_inertia[2];
    s.px += dt * s.vx; s.py += dt * s.vy; s.pz += dt * s.vz;
    // q <- q * exp(omega_b dt / 2), renormalised.  With h = |omega| dt / 2 the increment is
    // (omega * dt/2 * sin(h)/h, cos(h)); for h < 0.1 the even series in h^2 are exact to rounding
    // (no sqrt, no range reduction); a tumbling drone takes the sincos path.
    const R h2 = (s.wx * s.wx + s.wy * s.wy + s.wz * s.wz) * ((R)0.25 * dt * dt);
    R sinc, ch;
    if (h2 < (R)0.01) {
        sinc = (R)1 + h2 * ((R)(-1.0 / 6) + h2 * ((R)(1.0 / 120) + h2 * ((R)(-1.0 / 5040) + h2 * (R)(1.0 / 362880))));
        ch = (R)1 + h2 * ((R)-0.5 + h2 * ((R)(1.0 / 24) + h2 * ((R)(-1.0 / 720) + h2 * ((R)(1.0 / 40320) + h2 * (R)(-1.0 / 3628800)))));
    } else {
        const R h = sqrt_(h2);
        R sh;
        sincos_(h, &sh, &ch);
        sinc = sh / h;
    }
    const R k = (R)0.5 * dt * sinc;
    const R dx = s.wx * k, dy = s.wy * k, dz = s.wz * k, dw = ch;
    const R nx = w * dx + x * dw + y * dz - z * dy;
    const R ny = w * dy - x * dz + y * dw + z * dx;
    const R nzq = w * dz + x * dy - y * dx + z * dw;
    const R nw = w * dw - x * dx - y * dy - z * dz;
    const R inv = rsqrt_(nx * nx + ny * ny + nzq * nzq + nw * nw);
    s.qx = nx * inv; s.qy = ny * inv; s.qz = nzq * inv; s.qw = nw * inv;
    // static plane at z = -6 (entities_manager.py:121-125): inelastic clamp
    if (s.pz < P.ground_z) {
        s.pz = P.ground_z;
        if (s.vz < 0) s.vz = 0;
    }
}

// One physics substep of the reference loop (level4_simulation.py:84-98) in mode 6.
template <typename R, bool NOISE>
__device__ __forceinline__ void quad_substep(Drone<R>& s, const R sp[4], const QuadParams<R>& P,
                                             Imu<R>& imu, uint32_t k0, uint32_t k1, uint32_t env,
                                             uint32_t slot, uint32_t phys_step) {
    const Rot<R> M = quat_rot(s.qx, s.qy, s.qz, s.qw);
    const Wrench<R> W = quad_forces<R, NOISE>(s, sp, false, P, imu, k0, k1, env, slot, phys_step, M);
    quad_integrate<R>(s, W, P, M);
}

}  // namespace dc

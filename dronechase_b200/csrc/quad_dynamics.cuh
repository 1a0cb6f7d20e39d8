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
    // products the float32 substep uses as single constants (quad_substep_f32)
    R dt_im, dt_g, dt_iI[3], thrust_c, arm_c, km_c, half_dt, quarter_dt2;
};

// dc_config.quad (88 doubles, layout of oracle/dynamics.py QuadParams.flat()) -> device constants
template <typename R> __host__ __device__ constexpr QuadParams<R> make_quad(const double* f) {
    QuadParams<R> q{};
    const double mass = f[0], ix = f[1], iy = f[2], iz = f[3], arm = f[4], kf = f[5], km = f[6], tau = f[7];
    const double noise = f[8], max_rpm = f[9], drag_k = f[10], dt = f[11], pid_T = f[12], gyro = f[13];
    q.mass = (R)mass; q.inv_mass = (R)(1.0 / mass);
    q.inertia[0] = (R)ix; q.inertia[1] = (R)iy; q.inertia[2] = (R)iz;
    q.inv_inertia[0] = (R)(1.0 / ix); q.inv_inertia[1] = (R)(1.0 / iy); q.inv_inertia[2] = (R)(1.0 / iz);
    q.arm = (R)arm; q.kf = (R)kf; q.km = (R)km; q.dt_over_tau = (R)(dt / tau); q.noise_ratio = (R)noise;
    q.max_rpm = (R)max_rpm; q.drag_k = (R)drag_k; q.dt = (R)dt; q.pid_T = (R)pid_T; q.inv_pid_T = (R)(1.0 / pid_T);
    q.gravity = (R)f[14]; q.ground_z = (R)f[15]; q.gyro = gyro != 0.0;
    const double* g = f + 16;
    for (int p = 0; p < 6; ++p)
        for (int k = 0; k < 3; ++k) {
            q.kp[p][k] = (R)g[p * 12 + k]; q.ki[p][k] = (R)g[p * 12 + 3 + k];
            q.kd[p][k] = (R)g[p * 12 + 6 + k]; q.lim[p][k] = (R)g[p * 12 + 9 + k];
            q.kiT[p][k] = (R)(g[p * 12 + 3 + k] * pid_T); q.kdiT[p][k] = (R)(g[p * 12 + 6 + k] / pid_T);
        }
    q.dt_im = (R)(dt / mass); q.dt_g = (R)(dt * f[14]);
    q.dt_iI[0] = (R)(dt / ix); q.dt_iI[1] = (R)(dt / iy); q.dt_iI[2] = (R)(dt / iz);
    q.thrust_c = (R)(kf * max_rpm * max_rpm); q.arm_c = (R)(arm * kf * max_rpm * max_rpm); q.km_c = (R)(km * max_rpm * max_rpm);
    q.half_dt = (R)(0.5 * dt); q.quarter_dt2 = (R)(0.25 * dt * dt);
    return q;
}

// The drone every preset flies: dronechase_b200/config.py CF2X through quad_param_vector() (PyFlyt's cf2x.yaml schema).
// When dc_config.quad equals this table (noise ratio and ground height aside, which stay run-time values) the float32
// dynamics run an instantiation with the model folded into immediates: no constant loads in the substep loop, and the
// controllers whose ki / kd are zero lose those terms at compile time.
constexpr double CF2X_FLAT[88] = {
    0.027, 1.4e-05, 1.4e-05, 2.17e-05, 0.028, 3.16e-10, 7.94e-12, 0.01, 0.02, 21579.26219688767, 0.0024500000000000004,
    1.0 / 240, 1.0 / 120, 0.0, -9.81, -6.0,
    0.008, 0.008, 0.01, 2.5e-07, 2.5e-07, 0.00013, 0.0001, 0.0001, 0.0, 1.0, 1.0, 1.0,       // ang_vel  kp ki kd lim
    2.0, 2.0, 2.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 3.0, 3.0, 3.0,                              // ang_pos
    0.8, 0.8, 0.0, 0.3, 0.3, 0.0, 0.5, 0.5, 0.0, 0.4, 0.4, 0.0,                              // lin_vel
    0.15, 0.0, 0.0, 1.0, 0.0, 0.0, 0.015, 0.0, 0.0, 1.0, 0.0, 0.0,                           // z_vel
    1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 2.0, 2.0, 0.0,                              // lin_pos
    1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0};                             // z_pos
template <typename R> struct Cf2x { static constexpr QuadParams<R> v = make_quad<R>(CF2X_FLAT); };
inline bool quad_is_builtin(const double* f) {
    for (int i = 0; i < 88; ++i)
        if (i != 8 && i != 15 && f[i] != CF2X_FLAT[i]) return false;
    return true;
}

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

// ================================================================================================
// float32 product path: one substep of the mode-6 cascade (every family but stage01), written for instruction count --
// dyn_kernel is bound by instruction issue, not by memory.  Same physics as quad_forces +
// quad_integrate; what differs is rounding order only:
//   * the rotation matrix is formed from the doubled quaternion (x2 = x + x, ...): 20 instructions, kept live for the
//     euler angles, the body velocities and the force rotation;
//   * roll / yaw take R22 / R00 for Bullet's w^2 - x^2 - y^2 + z^2 / w^2 + x^2 - y^2 - z^2 (equal for a unit quaternion);
//   * the out-of-envelope cases (|pitch| > 30 deg, |roll| > 24 deg, gimbal lock) leave the loop through one test into a
//     call (euler_general) instead of three reconvergence regions; motor saturation likewise (mix_saturate);
//   * dt / m, dt / I, kf rpm_max^2, arm kf rpm_max^2, km rpm_max^2 are single constants;
//   * BUILTIN: the cf2x model as immediates (Cf2x<float>), controllers with ki = 0 / kd = 0 shortened at compile time;
//   * WANT_IMU = false (every substep but the last): the IMU record is not materialised -- only the record of the LAST
//     substep is ever read (level4_simulation.py:84-98: the snapshot precedes the last stepSimulation).
// ================================================================================================
// returns (roll, pitch, sin yaw, cos yaw) by value: results passed through pointers would pin the callers' variables to
// local memory on the fast path too
__device__ __noinline__ float4 euler_general(float x, float y, float z, float w, float r20, float r21, float r10) {
    const float sarg = -r20;
    float roll, pitch, sy, cy;
    if (sarg <= -0.99999f || sarg >= 0.99999f) {          // gimbal-lock branch of btQuaternion::getEulerZYX
        pitch = sarg < 0 ? -1.5707963267948966f : 1.5707963267948966f; roll = 0;
        const float yaw = sarg < 0 ? 2 * atan2f(x, -y) : 2 * atan2f(-x, y);
        sincosf(yaw, &sy, &cy);
    } else {
        pitch = asinf(sarg);
        roll = atan2f(r21, w * w - x * x - y * y + z * z);
        const float ys = r10, yc = w * w + x * x - y * y - z * z;
        const float h2 = ys * ys + yc * yc;
        const float ih = h2 > 0 ? rsqrtf(h2) : 0;
        sy = ys * ih; cy = h2 > 0 ? yc * ih : 1.0f;
    }
    return make_float4(roll, pitch, sy, cy);
}

// PyFlyt's motor mix saturation handling (out of line: reached by a few percent of the substeps)
__device__ __noinline__ float4 mix_saturate(float p0, float p1, float p2, float p3, float high, float low) {
    float pwm[4] = {p0, p1, p2, p3};
    if (high > 1.0f) {
        const float ih = mufu_rcp(high);
#pragma unroll
        for (int m = 0; m < 4; ++m) pwm[m] *= ih;
        low = fminf(fminf(pwm[0], pwm[1]), fminf(pwm[2], pwm[3]));
    }
    if (low < 0.05f) {
        const float kl = (0.05f - low) * mufu_rcp(1.0f - low);
#pragma unroll
        for (int m = 0; m < 4; ++m) pwm[m] = fmaf(1.0f - pwm[m], kl, pwm[m]);
    }
    return make_float4(pwm[0], pwm[1], pwm[2], pwm[3]);
}

// (sin(h) / h, cos(h)) for h^2 >= 0.01: a tumbling drone
__device__ __noinline__ float2 sinc_cos_general(float h2) {
    const float h = sqrtf(h2);
    float sh, ch;
    sincosf(h, &sh, &ch);
    return make_float2(sh / h, ch);
}

template <bool FOLD>
__device__ __forceinline__ float pid_f32(float& integ, float& prev, float kp, float kiT, float kdiT, float lim, float state, float sp) {
    const float err = sp - state;
    float out;
    if (!FOLD || kiT != 0.0f) { integ = clamp_(fmaf(kiT, err, integ), -lim, lim); out = fmaf(kp, err, integ); }
    else out = kp * err;                                   // ki = 0: the integrator never leaves 0
    if (!FOLD || kdiT != 0.0f) { out = fmaf(kdiT, err - prev, out); prev = err; }     // kd = 0: the previous error is never read
    return clamp_(out, -lim, lim);
}

template <bool NOISE, bool BUILTIN, bool WANT_IMU>
__device__ __forceinline__ void quad_substep_f32(Drone<float>& s, const float sp[4], const QuadParams<float>& Prt, Imu<float>& imu,
                                                 const uint32_t* rk, uint32_t env, uint32_t slot, uint32_t phys_step) {
#define QC(m) (BUILTIN ? Cf2x<float>::v.m : Prt.m)
#define QPID(i, j) QC(kp[i][j]), QC(kiT[i][j]), QC(kdiT[i][j]), QC(lim[i][j])
    const float x = s.qx, y = s.qy, z = s.qz, w = s.qw;
    // ---- body->world rotation ----
    const float x2 = x + x, y2 = y + y, z2 = z + z;
    const float xx = x2 * x, yy = y2 * y, zz = z2 * z, xy = x2 * y, xz = x2 * z, yz = y2 * z;
    const float ux = 1.0f - xx, uy = 1.0f - yy;
    const float r00 = uy - zz, r11 = ux - zz, r22 = ux - yy;
    const float r01 = fmaf(-z2, w, xy), r10 = fmaf(z2, w, xy);
    const float r02 = fmaf(y2, w, xz), r20 = fmaf(-y2, w, xz);
    const float r12 = fmaf(-x2, w, yz), r21 = fmaf(x2, w, yz);
    // ---- QuadX.update_state ----
    const float ub = fmaf(r20, s.vz, fmaf(r10, s.vy, r00 * s.vx));     // R^T v
    const float vb = fmaf(r21, s.vz, fmaf(r11, s.vy, r01 * s.vx));
    const float wb = fmaf(r22, s.vz, fmaf(r12, s.vy, r02 * s.vx));
    const float sarg = -r20;
    float roll, pitch, sy, cy;
    if (fabsf(sarg) <= 0.5f && r22 > 1e-6f && fabsf(r21) <= 0.45f * r22) {
        {   // asin: x + x^3 P(x^2), |x| <= 0.5 (asin_small)
            const float t = sarg * sarg;
            float p = 0.042095039f;
            p = fmaf(p, t, 0.024220509f); p = fmaf(p, t, 0.045462500f); p = fmaf(p, t, 0.074953534f); p = fmaf(p, t, 0.16666752f);
            pitch = fmaf(p * t, sarg, sarg);
        }
        {   // atan2(r21, r22), |r21| <= 0.45 r22 (atan2_small)
            const float t = r21 * mufu_rcp(r22), t2 = t * t;
            float p = -1.0f / 15.0f;
            p = fmaf(p, t2, 1.0f / 13.0f); p = fmaf(p, t2, -1.0f / 11.0f); p = fmaf(p, t2, 1.0f / 9.0f);
            p = fmaf(p, t2, -1.0f / 7.0f); p = fmaf(p, t2, 1.0f / 5.0f); p = fmaf(p, t2, -1.0f / 3.0f);
            roll = fmaf(p * t2, t, t);
        }
        const float ih = mufu_rsq(fmaf(r10, r10, r00 * r00));      // cos^2(pitch) >= 0.75 here
        sy = r10 * ih; cy = r00 * ih;
    } else {
        const float4 e = euler_general(x, y, z, w, r20, r21, r10);
        roll = e.x; pitch = e.y; sy = e.z; cy = e.w;
    }
    if (WANT_IMU) {
        imu.px = s.px; imu.py = s.py; imu.pz = s.pz;
        imu.roll = roll; imu.pitch = pitch;
        imu.ub = ub; imu.vb = vb; imu.wb = wb;
        imu.p = s.wx; imu.q = s.wy; imu.r = s.wz;
        imu.qx = x; imu.qy = y; imu.qz = z; imu.qw = w;
    }
    // ---- QuadX.update_control: the mode 6 cascade ----
    float* pid = s.pid;
    const float u_cmd = fmaf(sy, sp[1], cy * sp[0]);
    const float v_cmd = fmaf(cy, sp[1], -sy * sp[0]);
    const float o0 = pid_f32<BUILTIN>(pid[12], pid[14], QPID(2, 0), ub, u_cmd);
    const float o1 = pid_f32<BUILTIN>(pid[13], pid[15], QPID(2, 1), vb, v_cmd);
    const float p_cmd = pid_f32<BUILTIN>(pid[6], pid[9], QPID(1, 0), roll, -o1);
    const float q_cmd = pid_f32<BUILTIN>(pid[7], pid[10], QPID(1, 1), pitch, o0);
    const float tx = pid_f32<BUILTIN>(pid[0], pid[3], QPID(0, 0), s.wx, p_cmd);
    const float ty = pid_f32<BUILTIN>(pid[1], pid[4], QPID(0, 1), s.wy, q_cmd);
    const float tz = pid_f32<BUILTIN>(pid[2], pid[5], QPID(0, 2), s.wz, sp[2]);
    const float th = __saturatef(pid_f32<BUILTIN>(pid[16], pid[17], QPID(3, 0), wb, sp[3]));
    // motor mixing; saturation handling out of line
    const float ma = tz + th, mb = th - tz, mc = tx + ty, md = tx - ty;
    float pwm[4] = {ma - mc, ma + mc, mb - md, mb + md};
    {
        const float high = fmaxf(fmaxf(pwm[0], pwm[1]), fmaxf(pwm[2], pwm[3]));
        const float low = fminf(fminf(pwm[0], pwm[1]), fminf(pwm[2], pwm[3]));
        if (high > 1.0f || low < 0.05f) {
            const float4 m4 = mix_saturate(pwm[0], pwm[1], pwm[2], pwm[3], high, low);
            pwm[0] = m4.x; pwm[1] = m4.y; pwm[2] = m4.z; pwm[3] = m4.w;
        }
    }
    // ---- QuadX.update_physics: Motors + BoringBodies ----
    float nz[4] = {0, 0, 0, 0};
    if (NOISE) {
        const uint4 r = philox4x32_10_rk(phys_step, slot * 256u + (uint32_t)STREAM_MOTOR, env, 0u, rk);
        const float inv24 = (float)(1.0 / 16777216.0);
        const float u1 = fmaf((float)(r.x >> 8), inv24, inv24), u3 = fmaf((float)(r.z >> 8), inv24, inv24);   // (k + 1) / 2^24, exact
        const float r1 = mufu_sqrt(-1.3862943611198906f * mufu_lg2(u1)), r2 = mufu_sqrt(-1.3862943611198906f * mufu_lg2(u3));
        float sn, cs;
        bm_angle(r.y >> 8, &sn, &cs); nz[0] = r1 * cs; nz[1] = r1 * sn;
        bm_angle(r.w >> 8, &sn, &cs); nz[2] = r2 * cs; nz[3] = r2 * sn;
    }
    float tt[4];
    const float dtt = QC(dt_over_tau);
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        float t = fmaf(dtt, pwm[m] - s.thr[m], s.thr[m]);
        if (NOISE) t *= fmaf(nz[m], Prt.noise_ratio, 1.0f);
        s.thr[m] = t;
        tt[m] = t * t;
    }
    // propeller signs of the motor map: m0 (+,-) m1 (-,+) m2 (-,-) m3 (+,+); m0, m1 spin one way, m2, m3 the other
    const float s01 = tt[0] + tt[1], s23 = tt[2] + tt[3], d10 = tt[1] - tt[0], d32 = tt[3] - tt[2];
    float tau_x = QC(arm_c) * (d10 + d32);
    float tau_y = QC(arm_c) * (d10 - d32);
    float tau_z = QC(km_c) * (s01 - s23);
    const float dk = QC(drag_k);
    const float fbx = (-dk * ub) * fabsf(ub);
    const float fby = (-dk * vb) * fabsf(vb);
    const float fbz = fmaf(QC(thrust_c), s01 + s23, (-dk * wb) * fabsf(wb));
    // ---- stepSimulation: semi-implicit Euler of one free rigid body ----
    const float dim = QC(dt_im), dt = QC(dt);
    s.vx = fmaf(dim, fmaf(r02, fbz, fmaf(r01, fby, r00 * fbx)), s.vx);
    s.vy = fmaf(dim, fmaf(r12, fbz, fmaf(r11, fby, r10 * fbx)), s.vy);
    s.vz = fmaf(dim, fmaf(r22, fbz, fmaf(r21, fby, r20 * fbx)), s.vz) + QC(dt_g);
    if (QC(gyro)) {
        const float Ix = QC(inertia[0]) * s.wx, Iy = QC(inertia[1]) * s.wy, Iz = QC(inertia[2]) * s.wz;
        tau_x -= s.wy * Iz - s.wz * Iy;
        tau_y -= s.wz * Ix - s.wx * Iz;
        tau_z -= s.wx * Iy - s.wy * Ix;
    }
    s.wx = fmaf(QC(dt_iI[0]), tau_x, s.wx);
    s.wy = fmaf(QC(dt_iI[1]), tau_y, s.wy);
    s.wz = fmaf(QC(dt_iI[2]), tau_z, s.wz);
    s.px = fmaf(dt, s.vx, s.px); s.py = fmaf(dt, s.vy, s.py); s.pz = fmaf(dt, s.vz, s.pz);
    // q <- q * exp(omega_b dt / 2), renormalised (see quad_integrate); float32 needs two series terms for h^2 < 0.01
    const float h2 = fmaf(s.wz, s.wz, fmaf(s.wy, s.wy, s.wx * s.wx)) * QC(quarter_dt2);
    float sinc, ch;
    if (h2 < 0.01f) {
        sinc = fmaf(h2, fmaf(h2, 1.0f / 120, -1.0f / 6), 1.0f);
        ch = fmaf(h2, fmaf(h2, fmaf(h2, -1.0f / 720, 1.0f / 24), -0.5f), 1.0f);
    } else {
        const float2 t = sinc_cos_general(h2);
        sinc = t.x; ch = t.y;
    }
    const float k = QC(half_dt) * sinc;
    const float dx = s.wx * k, dy = s.wy * k, dz = s.wz * k;
    const float nx = fmaf(-z, dy, fmaf(y, dz, fmaf(x, ch, w * dx)));
    const float ny = fmaf(z, dx, fmaf(y, ch, fmaf(-x, dz, w * dy)));
    const float nq = fmaf(z, ch, fmaf(-y, dx, fmaf(x, dy, w * dz)));
    const float nw = fmaf(-z, dz, fmaf(-y, dy, fmaf(-x, dx, w * ch)));
    const float inv = mufu_rsq(fmaf(nw, nw, fmaf(nq, nq, fmaf(ny, ny, nx * nx))));
    s.qx = nx * inv; s.qy = ny * inv; s.qz = nq * inv; s.qw = nw * inv;
    // static plane at z = -6 (entities_manager.py:121-125): inelastic clamp
    if (s.pz < Prt.ground_z) {
        s.pz = Prt.ground_z;
        s.vz = fmaxf(s.vz, 0.0f);
    }
#undef QC
#undef QPID
}

}  // namespace dc

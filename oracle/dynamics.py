"""CPU oracle -- QuadX drone dynamics (TEST INFRASTRUCTURE, not the product).

numpy float64 restatement of the third-party arithmetic the reference calls on
its hot path but does not contain:

  * ``PyFlyt.core.drones.quadx.QuadX`` (pyflyt==0.11.1, poetry.lock:1432-1449):
    ``update_state`` / ``update_control`` (mode 6 and 7 PID cascade) /
    ``update_physics`` (first-order motor lag + multiplicative noise, thrust
    and reaction torque ~ rpm^2, per-axis quadratic body drag);
  * ``pybullet==3.2.7`` ``stepSimulation`` for one free rigid body (semi-implicit
    Euler, dt = 1/240, gravity -9.81) and its quaternion/euler utilities.

Reference call sites this follows (order of operations per physics substep):
  src/threatengage/environments/level4/components/simulation/level4_simulation.py:84-98
  src/core/entities/quadcopters/quadcopter.py:143-152, 379-413, 433-482, 543-549
  src/core/entities/quadcopters/components/sensors/imu.py:27-41

PARITY UNPINNED for this file: neither wheel is installed in the build
container or vendored under /root/reference, and the reference has no test
that pins a trajectory.  The constants below are the
published cf2x model as best recalled; they live in one dict with PyFlyt's
yaml schema so that the real ``cf2x.yaml`` can be dropped in.  The CUDA path is
checked against THIS restatement (tests/), not against PyFlyt itself.

Every function is vectorised over arbitrary leading batch dimensions so the
same code serves (a) the fake ``pybullet``/``PyFlyt`` backend used to execute
the reference's own game logic (oracle/refshim) and (b) the batched env oracle.
"""
from __future__ import annotations

import numpy as np

# ---------------------------------------------------------------------------
# cf2x model (PyFlyt yaml schema: motor_params / drag_params / control_params)
# ---------------------------------------------------------------------------
CF2X = {
    "mass": 0.027,
    "inertia": [1.4e-5, 1.4e-5, 2.17e-5],
    "arm": 0.028,  # |x| = |y| of each propeller in the body frame
    "motor_params": {
        "total_thrust": 0.5886,
        "thrust_coef": 3.16e-10,
        "torque_coef": 7.94e-12,
        "noise_ratio": 0.02,
        "tau": 0.01,
    },
    "drag_params": {"drag_coef_xyz": 1.0, "drag_area_xyz": 0.004},
    "control_params": {
        "ang_vel": {"kp": [8e-3, 8e-3, 1e-2], "ki": [2.5e-7, 2.5e-7, 1.3e-4],
                    "kd": [1e-4, 1e-4, 0.0], "lim": [1.0, 1.0, 1.0]},
        "ang_pos": {"kp": [2.0, 2.0, 2.0], "ki": [0.0, 0.0, 0.0],
                    "kd": [0.0, 0.0, 0.0], "lim": [3.0, 3.0, 3.0]},
        "lin_vel": {"kp": [0.8, 0.8], "ki": [0.3, 0.3], "kd": [0.5, 0.5],
                    "lim": [0.4, 0.4]},
        "lin_pos": {"kp": [1.0, 1.0], "ki": [0.0, 0.0], "kd": [0.0, 0.0],
                    "lim": [2.0, 2.0]},
        "z_pos": {"kp": 1.0, "ki": 0.0, "kd": 0.0, "lim": 1.0},
        "z_vel": {"kp": 0.15, "ki": 1.0, "kd": 0.015, "lim": 1.0},
    },
}

GRAVITY = -9.81            # level4_simulation.py:70
PHYSICS_HZ = 240           # level4_simulation.py:29
CONTROL_HZ = 120           # QuadX default control_hz; PID period = 1/120
RHO_AIR = 1.225
GROUND_Z = -6.0            # entities_manager.py:121-125

# motor order/mixing of PyFlyt QuadX: rows = motors, cols = (roll, pitch, yaw, thrust)
MOTOR_MAP = np.array([[-1.0, -1.0, +1.0, +1.0],
                      [+1.0, +1.0, +1.0, +1.0],
                      [-1.0, +1.0, -1.0, +1.0],
                      [+1.0, -1.0, -1.0, +1.0]])
# propeller x/y signs consistent with MOTOR_MAP (torque = r x F, F along +z)
MOTOR_X = np.array([+1.0, -1.0, -1.0, +1.0])
MOTOR_Y = np.array([-1.0, +1.0, -1.0, +1.0])
MOTOR_YAW = np.array([+1.0, +1.0, -1.0, -1.0])

# PID state layout (24 words per drone), shared with the CUDA kernels
PID_SLOTS = {
    "ang_vel": (0, 3), "ang_pos": (6, 3), "lin_vel": (12, 2),
    "z_vel": (16, 1), "lin_pos": (18, 2), "z_pos": (22, 1),
}
PID_WORDS = 24


class QuadParams:
    """Flattened numeric view of a cf2x-style model dict."""

    def __init__(self, model: dict = CF2X, noise_ratio: float | None = None,
                 gyro_term: bool = False, ground_z: float = GROUND_Z):
        self.ground_z = float(ground_z)       # level2/level3 have no plane: pass a very low value
        self.mass = float(model["mass"])
        self.inertia = np.asarray(model["inertia"], dtype=np.float64)
        self.arm = float(model["arm"])
        mp = model["motor_params"]
        self.thrust_coef = float(mp["thrust_coef"])
        self.torque_coef = float(mp["torque_coef"])
        self.tau = float(mp["tau"])
        self.noise_ratio = float(mp["noise_ratio"] if noise_ratio is None else noise_ratio)
        self.max_rpm = float(np.sqrt(mp["total_thrust"] / (4.0 * mp["thrust_coef"])))
        dp = model["drag_params"]
        self.drag_k = 0.5 * RHO_AIR * float(dp["drag_coef_xyz"]) * float(dp["drag_area_xyz"])
        self.gains = {}
        for name, g in model["control_params"].items():
            n = PID_SLOTS[name][1]
            self.gains[name] = tuple(
                np.broadcast_to(np.asarray(g[k], dtype=np.float64), (n,)).copy()
                for k in ("kp", "ki", "kd", "lim"))
        self.dt = 1.0 / PHYSICS_HZ
        self.pid_period = 1.0 / CONTROL_HZ
        self.gyro_term = bool(gyro_term)

    def flat(self) -> np.ndarray:
        """Parameter vector handed to the C ABI (see include/dronechase_b200.h)."""
        out = [self.mass, *self.inertia, self.arm, self.thrust_coef, self.torque_coef,
               self.tau, self.noise_ratio, self.max_rpm, self.drag_k, self.dt,
               self.pid_period, float(self.gyro_term), GRAVITY, self.ground_z]
        for name in ("ang_vel", "ang_pos", "lin_vel", "z_vel", "lin_pos", "z_pos"):
            kp, ki, kd, lim = self.gains[name]
            n = PID_SLOTS[name][1]
            for arr in (kp, ki, kd, lim):
                out.extend(list(arr) + [0.0] * (3 - n))
        return np.asarray(out, dtype=np.float64)


# ---------------------------------------------------------------------------
# quaternion helpers -- PyBullet conventions, quaternions are [x, y, z, w]
# ---------------------------------------------------------------------------
def quat_from_euler(e):
    """p.getQuaternionFromEuler([roll, pitch, yaw]) (quadcopter.py:434, imu.py:38)."""
    e = np.asarray(e, dtype=np.float64)
    hr, hp, hy = 0.5 * e[..., 0], 0.5 * e[..., 1], 0.5 * e[..., 2]
    cr, sr, cp, sp, cy, sy = np.cos(hr), np.sin(hr), np.cos(hp), np.sin(hp), np.cos(hy), np.sin(hy)
    return np.stack([sr * cp * cy - cr * sp * sy,
                     cr * sp * cy + sr * cp * sy,
                     cr * cp * sy - sr * sp * cy,
                     cr * cp * cy + sr * sp * sy], axis=-1)


def euler_from_quat(q):
    """p.getEulerFromQuaternion (btQuaternion::getEulerZYX): returns [roll, pitch, yaw]."""
    q = np.asarray(q, dtype=np.float64)
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    sarg = -2.0 * (x * z - w * y)
    roll = np.arctan2(2.0 * (y * z + w * x), w * w - x * x - y * y + z * z)
    pitch = np.arcsin(np.clip(sarg, -1.0, 1.0))
    yaw = np.arctan2(2.0 * (x * y + w * z), w * w + x * x - y * y - z * z)
    lo, hi = sarg <= -0.99999, sarg >= 0.99999
    roll = np.where(lo | hi, 0.0, roll)
    pitch = np.where(lo, -0.5 * np.pi, np.where(hi, 0.5 * np.pi, pitch))
    yaw = np.where(lo, 2.0 * np.arctan2(x, -y), np.where(hi, 2.0 * np.arctan2(-x, y), yaw))
    return np.stack([roll, pitch, yaw], axis=-1)


def rot_from_quat(q):
    """Rotation matrix body->world, rows as p.getMatrixFromQuaternion lists them."""
    q = np.asarray(q, dtype=np.float64)
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z); R[..., 0, 1] = 2 * (x * y - w * z); R[..., 0, 2] = 2 * (x * z + w * y)
    R[..., 1, 0] = 2 * (x * y + w * z); R[..., 1, 1] = 1 - 2 * (x * x + z * z); R[..., 1, 2] = 2 * (y * z - w * x)
    R[..., 2, 0] = 2 * (x * z - w * y); R[..., 2, 1] = 2 * (y * z + w * x); R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def rotate_vector(q, v):
    """p.rotateVector(q, v) = R(q) v (lidar_math.py:75,81)."""
    return np.einsum("...ij,...j->...i", rot_from_quat(q), np.asarray(v, dtype=np.float64))


def quat_mul(a, b):
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw,
                     aw * bw - ax * bx - ay * by - az * bz], axis=-1)


# ---------------------------------------------------------------------------
# QuadX.update_state  (imu.py:27-41 unpacks rows [ang_vel, euler, lin_vel, pos])
# ---------------------------------------------------------------------------
def imu_state(pos, quat, vel_w, omega_b):
    """State as the reference's IMU sees it: body-frame velocities, euler, pos."""
    R = rot_from_quat(quat)
    vel_b = np.einsum("...ji,...j->...i", R, vel_w)          # R^T v
    return {"position": np.array(pos, dtype=np.float64, copy=True),
            "attitude": euler_from_quat(quat),
            "velocity": vel_b,
            "angular_rate": np.array(omega_b, dtype=np.float64, copy=True),
            "quaternion": np.array(quat, dtype=np.float64, copy=True)}


# ---------------------------------------------------------------------------
# PyFlyt PID + QuadX.update_control
# ---------------------------------------------------------------------------
def _pid(pid, name, gains, state, setpoint, period, sel=slice(None)):
    """PyFlyt PID.step: I clipped, output clipped; (I, prev_err) live in pid[...,24]."""
    off, n = PID_SLOTS[name]
    kp, ki, kd, lim = (g[sel] for g in gains[name])
    idx_i = np.arange(off, off + n)[sel]
    idx_e = np.arange(off + 3, off + 3 + n)[sel] if n == 3 else np.arange(off + n, off + 2 * n)[sel]
    err = setpoint - state
    integ = np.clip(pid[..., idx_i] + ki * err * period, -lim, lim)
    deriv = kd * (err - pid[..., idx_e]) / period
    pid[..., idx_i] = integ
    pid[..., idx_e] = err
    return np.clip(kp * err + integ + deriv, -lim, lim)


def control_update(pid, imu, setpoint, mode, prm: QuadParams):
    """QuadX.update_control for mode 6 (vx, vy, vr, vz) / 7 (x, y, r, z) -> pwm[4].

    ``pid`` (..., 24) is updated in place.  ``mode`` may be a scalar or an array
    broadcastable to the batch shape.
    """
    sp = np.asarray(setpoint, dtype=np.float64)
    mode = np.broadcast_to(np.asarray(mode), sp.shape[:-1])
    T = prm.pid_period
    eul, vel_b, pos, rate = imu["attitude"], imu["velocity"], imu["position"], imu["angular_rate"]
    a_xy = sp[..., 0:2].copy()
    a_r = sp[..., 2].copy()
    z_out = sp[..., 3].copy()
    is7 = mode == 7
    if np.any(is7):
        pid7 = pid.copy()
        v_cmd = _pid(pid7, "lin_pos", prm.gains, pos[..., 0:2], a_xy, T)
        vz_cmd = _pid(pid7, "z_pos", prm.gains, pos[..., 2:3], z_out[..., None], T)[..., 0]
        a_xy = np.where(is7[..., None], v_cmd, a_xy)
        z_out = np.where(is7, vz_cmd, z_out)
        pid[...] = np.where(is7[..., None], pid7, pid)
    # ground-frame velocity command -> yaw-aligned body frame
    c, s = np.cos(eul[..., 2]), np.sin(eul[..., 2])
    u_cmd = np.stack([c * a_xy[..., 0] + s * a_xy[..., 1],
                      -s * a_xy[..., 0] + c * a_xy[..., 1]], axis=-1)
    out = _pid(pid, "lin_vel", prm.gains, vel_b[..., 0:2], u_cmd, T)
    ang_cmd = np.stack([-out[..., 1], out[..., 0]], axis=-1)            # (roll, pitch)
    rate_rp = _pid(pid, "ang_pos", prm.gains, eul[..., 0:2], ang_cmd, T, sel=slice(0, 2))
    if np.any(is7):
        pid7 = pid.copy()
        yaw_rate = _pid(pid7, "ang_pos", prm.gains, eul[..., 2:3], a_r[..., None], T, sel=slice(2, 3))[..., 0]
        a_r = np.where(is7, yaw_rate, a_r)
        pid[...] = np.where(is7[..., None], pid7, pid)
    rate_cmd = np.concatenate([rate_rp, a_r[..., None]], axis=-1)
    torque = _pid(pid, "ang_vel", prm.gains, rate, rate_cmd, T)
    thrust = _pid(pid, "z_vel", prm.gains, vel_b[..., 2:3], z_out[..., None], T)[..., 0]
    thrust = np.clip(thrust, 0.0, 1.0)
    cmd = np.concatenate([torque, thrust[..., None]], axis=-1)
    pwm = np.einsum("mk,...k->...m", MOTOR_MAP, cmd)
    high = np.max(pwm, axis=-1, keepdims=True)
    pwm = np.where(high > 1.0, pwm / np.where(high > 1.0, high, 1.0), pwm)
    low = np.min(pwm, axis=-1, keepdims=True)
    lift = (1.0 - pwm) / np.where(low < 0.05, 1.0 - low, 1.0) * (0.05 - low)
    pwm = np.where(low < 0.05, pwm + lift, pwm)
    return pwm


# ---------------------------------------------------------------------------
# QuadX.update_physics (Motors + BoringBodies) and Bullet stepSimulation
# ---------------------------------------------------------------------------
def actuate(throttle, pwm, vel_b, noise, prm: QuadParams):
    """Motor lag+noise, thrust/torque, drag.  Returns (throttle', F_body, tau_body)."""
    throttle = throttle + (prm.dt / prm.tau) * (pwm - throttle)
    throttle = throttle + noise * throttle * prm.noise_ratio
    rpm = throttle * prm.max_rpm
    thrust = prm.thrust_coef * rpm * rpm                      # (...,4)
    react = prm.torque_coef * rpm * rpm
    fz = np.sum(thrust, axis=-1)
    tau = np.stack([prm.arm * np.sum(MOTOR_Y * thrust, axis=-1),
                    -prm.arm * np.sum(MOTOR_X * thrust, axis=-1),
                    np.sum(MOTOR_YAW * react, axis=-1)], axis=-1)
    drag = -np.sign(vel_b) * prm.drag_k * vel_b * vel_b
    force = drag.copy()
    force[..., 2] += fz
    return throttle, force, tau


def rigid_body_step(pos, quat, vel_w, omega_b, force_b, tau_b, prm: QuadParams):
    """One stepSimulation for a free rigid body with a static plane at z = -6."""
    dt = prm.dt
    R = rot_from_quat(quat)
    acc = np.einsum("...ij,...j->...i", R, force_b) / prm.mass
    acc[..., 2] += GRAVITY
    vel_w = vel_w + dt * acc
    if prm.gyro_term:
        Iw = prm.inertia * omega_b
        tau_b = tau_b - np.cross(omega_b, Iw)
    omega_b = omega_b + dt * tau_b / prm.inertia
    pos = pos + dt * vel_w
    # orientation: q <- q * exp(omega_b dt / 2)
    th = np.linalg.norm(omega_b, axis=-1) * dt
    half = 0.5 * th
    k = np.where(th > 1e-12, np.sin(half) / np.where(th > 1e-12, th, 1.0) * dt, 0.5 * dt)
    dq = np.concatenate([omega_b * k[..., None], np.cos(half)[..., None]], axis=-1)
    quat = quat_mul(quat, dq)
    quat = quat / np.linalg.norm(quat, axis=-1, keepdims=True)
    # plane: inelastic clamp (only matters near z=-6 where the tasks terminate)
    below = pos[..., 2] < prm.ground_z
    pos = pos.copy(); vel_w = vel_w.copy()
    pos[..., 2] = np.where(below, prm.ground_z, pos[..., 2])
    vel_w[..., 2] = np.where(below & (vel_w[..., 2] < 0.0), 0.0, vel_w[..., 2])
    return pos, quat, vel_w, omega_b


def command_to_setpoint(command):
    """Quadcopter.convert_command_to_setpoint (quadcopter.py:379-396)."""
    command = np.asarray(command, dtype=np.float64)
    raw = command[..., 0:3]
    n = np.linalg.norm(raw, axis=-1, keepdims=True)
    direction = raw / np.where(n > 0, n, 1.0)
    v = command[..., 3:4] * direction
    return np.stack([v[..., 0], v[..., 1], np.zeros_like(v[..., 0]), v[..., 2]], axis=-1)

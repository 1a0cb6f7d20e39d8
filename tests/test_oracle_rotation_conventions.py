"""the rotation / Euler conventions `oracle/dynamics.py` restates for PyBullet (quaternions [x, y, z, w],
getQuaternionFromEuler / getEulerFromQuaternion = roll-pitch-yaw about the FIXED x, y, z axes, getMatrixFromQuaternion = body -> world,
the IMU's body-frame velocity R^T v, the integrator's q <- q * exp(omega_b dt / 2)) checked against an INDEPENDENT published
implementation, `scipy.spatial.transform.Rotation` (pybullet itself is not installable here: profiles/r2_ref_probe.log).
This does not pin the third-party dynamics; it removes "conventions recalled from memory" as a source of error."""
import numpy as np
import pytest

from oracle import dynamics as dy

Rotation = pytest.importorskip("scipy.spatial.transform").Rotation


def _same_rotation(qa, qb, tol=1e-12):
    return np.minimum(np.abs(qa - qb).max(axis=-1), np.abs(qa + qb).max(axis=-1)).max() < tol


def test_euler_quaternion_matrix_conventions_match_scipy():
    rng = np.random.RandomState(0)
    e = np.stack([rng.uniform(-np.pi, np.pi, 500), rng.uniform(-1.5, 1.5, 500), rng.uniform(-np.pi, np.pi, 500)], axis=1)
    q = dy.quat_from_euler(e)
    ref = Rotation.from_euler("xyz", e)                  # lower case: extrinsic rotations about x, then y, then z
    assert _same_rotation(q, ref.as_quat())
    assert np.abs(dy.rot_from_quat(q) - ref.as_matrix()).max() < 1e-12
    assert np.abs(dy.euler_from_quat(q) - ref.as_euler("xyz")).max() < 1e-9
    qr = rng.normal(size=(500, 4)); qr /= np.linalg.norm(qr, axis=1, keepdims=True)
    rr = Rotation.from_quat(qr)
    assert np.abs(dy.rot_from_quat(qr) - rr.as_matrix()).max() < 1e-12
    back = dy.euler_from_quat(qr)
    assert _same_rotation(dy.quat_from_euler(back), qr, 1e-9) and np.abs(back - rr.as_euler("xyz")).max() < 1e-8
    v = rng.normal(size=(500, 3))
    assert np.abs(dy.rotate_vector(qr, v) - rr.apply(v)).max() < 1e-12
    # Hamilton product in [x, y, z, w]: (a * b) rotates by b first, then a
    qb = rng.normal(size=(500, 4)); qb /= np.linalg.norm(qb, axis=1, keepdims=True)
    assert _same_rotation(dy.quat_mul(qr, qb), (rr * Rotation.from_quat(qb)).as_quat())


def test_imu_body_frame_and_integrator_orientation_update():
    rng = np.random.RandomState(1)
    q = rng.normal(size=(64, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    vel_w, omega_b = rng.normal(size=(64, 3)), rng.normal(size=(64, 3)) * 3
    imu = dy.imu_state(np.zeros((64, 3)), q, vel_w, omega_b)
    r = Rotation.from_quat(q)
    assert np.abs(imu["velocity"] - r.inv().apply(vel_w)).max() < 1e-12          # world velocity seen from the body
    assert np.abs(imu["attitude"] - r.as_euler("xyz")).max() < 1e-8
    # one torque-free, force-free physics step: the orientation advances by the body-frame rotation vector omega_b * dt
    prm = dy.QuadParams(gyro_term=False, ground_z=-1e9)
    pos, q1, v1, w1 = dy.rigid_body_step(np.zeros((64, 3)), q, np.zeros((64, 3)), omega_b, np.zeros((64, 3)), np.zeros((64, 3)), prm)
    assert np.abs(w1 - omega_b).max() < 1e-15
    assert _same_rotation(q1, (r * Rotation.from_rotvec(omega_b * prm.dt)).as_quat(), 1e-12)
    assert np.abs(v1[:, 2] - dy.GRAVITY * prm.dt).max() < 1e-15 and np.abs(pos[:, 2] - dy.GRAVITY * prm.dt ** 2).max() < 1e-15   # v first, then x


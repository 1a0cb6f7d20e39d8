"""CPU oracle -- stand-in physics backend for executing the reference's OWN game logic.

TEST INFRASTRUCTURE, build-container only.  The reference (/root/reference, pure
Python) cannot run here because ``pybullet``, ``PyFlyt``, ``gymnasium`` and
``pynput`` are not installed.  ``install()`` registers minimal
stand-ins for those four packages in ``sys.modules`` so that the reference's
environments, tasks, navigators, offset handler, gun and LiDAR classes import and
run UNMODIFIED from /root/reference/src.  The stand-in physics is
``oracle/dynamics.py`` (the float64 restatement of QuadX + one rigid body), so a
trajectory recorded this way pins the reference's *logic* (engagement, waves,
reward, termination, LiDAR, observation) on top of the restated dynamics.

Nothing here travels to the GPU box as a dependency: the vectors it produces are
committed under tests/golden/ by oracle/make_golden.py.
"""
from __future__ import annotations

import sys
import types

import numpy as np

from . import dynamics as dy

REFERENCE_SRC = "/root/reference/src"


class Hooks:
    """Injection points for the 'randomness as data' protocol (oracle/philox.py)."""
    prm = dy.QuadParams()
    # motor_noise(drone_index:int) -> 4 normals for the substep being integrated
    motor_noise = staticmethod(lambda drone_index: np.zeros(4))


class _Body:
    def __init__(self, pos, quat, mass):
        self.pos = np.array(pos, dtype=np.float64)
        self.quat = np.array(quat, dtype=np.float64)
        self.vel = np.zeros(3)
        self.omega_b = np.zeros(3)
        self.mass = float(mass)
        self.force_b = np.zeros(3)
        self.tau_b = np.zeros(3)


class BulletClient:
    """Subset of pybullet_utils.bullet_client.BulletClient the reference touches."""
    _n_clients = 0

    def __init__(self, connection_mode=None):
        self._client = BulletClient._n_clients
        BulletClient._n_clients += 1
        self.bodies = {}
        self._next_id = 0
        self.LINK_FRAME = 1

    # -- world management ---------------------------------------------------
    def resetSimulation(self):
        self.bodies = {}
        self._next_id = 0

    def setGravity(self, *a): pass
    def resetDebugVisualizerCamera(self, **k): pass
    def setAdditionalSearchPath(self, *a): pass
    def addUserDebugText(self, *a, **k): return 0
    def addUserDebugLine(self, *a, **k): return 0
    def removeUserDebugItem(self, *a, **k): pass
    def close(self): pass
    def disconnect(self): pass

    def _new_body(self, pos, quat, mass):
        bid = self._next_id
        self._next_id += 1
        self.bodies[bid] = _Body(pos, quat, mass)
        return bid

    def loadURDF(self, name, basePosition=(0, 0, 0), baseOrientation=(0, 0, 0, 1), **k):
        return self._new_body(basePosition, baseOrientation, 0.0)

    def removeBody(self, bid):
        self.bodies.pop(bid, None)

    # -- queries --------------------------------------------------------------
    def getDynamicsInfo(self, bid, link):
        return (self.bodies[bid].mass,)

    def getVisualShapeData(self, bid):
        return [(bid, -1, 5, (0.1, 0.1, 0.03))]

    def getNumJoints(self, bid):
        return 5

    def getBasePositionAndOrientation(self, bid):
        b = self.bodies[bid]
        return tuple(b.pos), tuple(b.quat)

    def getQuaternionFromEuler(self, e):
        return tuple(dy.quat_from_euler(np.asarray(e, dtype=np.float64)))

    # -- mutation ---------------------------------------------------------------
    def changeDynamics(self, bid, link, mass=None, **k):
        if mass is not None and link == -1:
            self.bodies[bid].mass = float(mass)

    def resetBaseVelocity(self, bid, linearVelocity=(0, 0, 0), angularVelocity=(0, 0, 0)):
        b = self.bodies[bid]
        b.vel = np.array(linearVelocity, dtype=np.float64)
        b.omega_b = np.array(angularVelocity, dtype=np.float64)

    def resetBasePositionAndOrientation(self, bid, pos, quat):
        b = self.bodies[bid]
        b.pos = np.array(pos, dtype=np.float64)
        b.quat = np.array(quat, dtype=np.float64)
        b.vel = np.zeros(3)
        b.omega_b = np.zeros(3)

    def stepSimulation(self):
        prm = Hooks.prm
        for b in self.bodies.values():
            if b.mass > 0.0:
                b.pos, b.quat, b.vel, b.omega_b = dy.rigid_body_step(
                    b.pos, b.quat, b.vel, b.omega_b, b.force_b, b.tau_b, prm)
            b.force_b = np.zeros(3)
            b.tau_b = np.zeros(3)


class _Resettable:
    def __init__(self, fn): self._fn = fn
    def reset(self): self._fn()


class QuadX:
    """Subset of PyFlyt.core.drones.quadx.QuadX the reference touches."""
    _order = []          # creation order -> drone slot index (for the noise stream)

    def __init__(self, p, start_pos, start_orn, control_hz=120, physics_hz=240,
                 np_random=None, **k):
        self.p = p
        self.Id = p._new_body(start_pos, dy.quat_from_euler(np.asarray(start_orn, float)),
                              Hooks.prm.mass)
        self.index = len(QuadX._order)
        QuadX._order.append(self.Id)
        self.control_period = 1.0 / control_hz
        self.physics_control_ratio = int(physics_hz // control_hz)
        self.mode = 0
        self.setpoint = np.zeros(4)
        self.pwm = np.zeros(4)
        self.throttle = np.zeros(4)
        self.pid = np.zeros(dy.PID_WORDS)
        self.state = np.zeros((4, 3))
        self._imu = None
        self.body = _Resettable(lambda: None)
        self.motors = _Resettable(self._reset_motors)

    def _reset_motors(self):
        self.throttle = np.zeros(4)

    def reset(self):
        self.setpoint = np.zeros(4)
        self.pwm = np.zeros(4)
        self._reset_motors()
        self.update_state()

    def set_mode(self, mode):
        self.mode = int(mode)
        self.pid = np.zeros(dy.PID_WORDS)
        self.setpoint = np.zeros(4)
        if mode == 7:
            b = self.p.bodies[self.Id]
            e = dy.euler_from_quat(b.quat)
            self.setpoint = np.array([b.pos[0], b.pos[1], e[2], b.pos[2]])

    def disable_artificial_damping(self): pass
    def update_last(self): pass

    def update_state(self):
        b = self.p.bodies[self.Id]
        self._imu = dy.imu_state(b.pos, b.quat, b.vel, b.omega_b)
        self.state = np.stack([self._imu["angular_rate"], self._imu["attitude"],
                               self._imu["velocity"], self._imu["position"]])

    def update_control(self):
        self.pwm = dy.control_update(self.pid, self._imu, np.asarray(self.setpoint, float),
                                     self.mode, Hooks.prm)

    def update_physics(self):
        b = self.p.bodies[self.Id]
        noise = np.asarray(Hooks.motor_noise(self.index), dtype=np.float64)
        self.throttle, f, t = dy.actuate(self.throttle, self.pwm, self._imu["velocity"],
                                         noise, Hooks.prm)
        b.force_b = b.force_b + f
        b.tau_b = b.tau_b + t


# ---------------------------------------------------------------------------
# gymnasium / pynput stand-ins
# ---------------------------------------------------------------------------
class _Env:
    metadata = {}
    def reset(self, seed=None, options=None): raise NotImplementedError
    def step(self, action): raise NotImplementedError
    def close(self): pass


class _Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape) if shape is not None else np.shape(low)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(self.dtype)


class _Dict:
    def __init__(self, spaces): self.spaces = dict(spaces)
    def __getitem__(self, k): return self.spaces[k]


class _MultiBinary:
    def __init__(self, n):
        self.n = n
        self.shape = tuple(n) if isinstance(n, (tuple, list)) else (n,)
        self.dtype = np.dtype(np.int8)


class _AnyKey:
    def __getattr__(self, name): return name


class _KeyCode:
    @staticmethod
    def from_char(c): return c


def install():
    """Register the stand-in packages and put the reference sources on sys.path."""
    pb = types.ModuleType("pybullet")
    pb.GUI, pb.DIRECT, pb.LINK_FRAME = 1, 2, 1
    pb.rotateVector = lambda q, v: tuple(dy.rotate_vector(np.asarray(q, float), np.asarray(v, float)))
    pb.getMatrixFromQuaternion = lambda q: tuple(dy.rot_from_quat(np.asarray(q, float)).reshape(9))
    pb.getQuaternionFromEuler = lambda e: tuple(dy.quat_from_euler(np.asarray(e, float)))
    pb.getEulerFromQuaternion = lambda q: tuple(dy.euler_from_quat(np.asarray(q, float)))
    pb.changeVisualShape = lambda *a, **k: None
    pb.setCollisionFilterGroupMask = lambda *a, **k: None
    pb.addUserDebugLine = lambda *a, **k: 0
    pbd = types.ModuleType("pybullet_data")
    pbd.getDataPath = lambda: ""
    pbu = types.ModuleType("pybullet_utils")
    pbc = types.ModuleType("pybullet_utils.bullet_client")
    pbc.BulletClient = BulletClient
    pbu.bullet_client = pbc

    names = ["PyFlyt", "PyFlyt.core", "PyFlyt.core.drones", "PyFlyt.core.drones.quadx"]
    mods = {n: types.ModuleType(n) for n in names}
    mods["PyFlyt.core.drones.quadx"].QuadX = QuadX
    mods["PyFlyt"].core = mods["PyFlyt.core"]
    mods["PyFlyt.core"].drones = mods["PyFlyt.core.drones"]
    mods["PyFlyt.core.drones"].quadx = mods["PyFlyt.core.drones.quadx"]

    gym = types.ModuleType("gymnasium")
    spaces = types.ModuleType("gymnasium.spaces")
    spaces.Box, spaces.Dict, spaces.MultiBinary = _Box, _Dict, _MultiBinary
    gym.Env, gym.spaces = _Env, spaces
    envs = types.ModuleType("gymnasium.envs")
    reg = types.ModuleType("gymnasium.envs.registration")
    reg.register = lambda *a, **k: None
    envs.registration = reg
    gym.envs = envs

    pyn = types.ModuleType("pynput")
    kb = types.ModuleType("pynput.keyboard")
    kb.Key, kb.KeyCode = _AnyKey(), _KeyCode
    pyn.keyboard = kb

    sb3 = types.ModuleType("stable_baselines3")
    sb3.PPO = type("PPO", (), {"load": staticmethod(lambda *a, **k: None)})
    sb3c = types.ModuleType("stable_baselines3.common")
    sb3.common = sb3c
    sys.modules.update({"stable_baselines3": sb3, "stable_baselines3.common": sb3c})

    sys.modules.update({"pybullet": pb, "pybullet_data": pbd, "pybullet_utils": pbu,
                        "pybullet_utils.bullet_client": pbc, "gymnasium": gym,
                        "gymnasium.spaces": spaces, "gymnasium.envs": envs,
                        "gymnasium.envs.registration": reg, "pynput": pyn,
                        "pynput.keyboard": kb, **mods})
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)


def fresh_singletons():
    """The reference keeps thread-local singletons (entities_manager.py:24-30,
    message_hub.py:10-16); drop them so a new env starts from a clean registry."""
    import threading
    from core.notification_system.message_hub import MessageHub
    MessageHub._thread_local_data = threading.local()
    for modname in ("threatengage.environments.level4.components.entities_management.entities_manager",
                    "threatsense.level5.components.entities_manager"):
        mod = sys.modules.get(modname)
        if mod is not None:
            mod.EntitiesManager._thread_local_data = threading.local()
    QuadX._order = []

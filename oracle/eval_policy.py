"""Deterministic stand-in for the SB3 PPO drivers of the level4 evaluation / exp05 tasks (TEST INFRASTRUCTURE).

The reference drives "nn" wingmen with ``PPO.load(path).predict(observation, deterministic=True)``
(evaluation_task.py:260-271,630-634; exp05_vFinal_task.py:252-260).  No trained model travels with the repository, so the
golden recordings, the oracle and the GPU tests all use this fixed function of the observation dict instead -- it reads every
key the real policy reads (``lidar`` (3,13,26), ``inertial_data`` (15,), ``last_action`` (4,)), so a wrong sphere, a wrong gun
state or a wrong shared ``last_action`` changes the action and with it the trajectory.

    nearest munition cell of the sphere (flag channel 0.2) -> unit vector of the cell centre, body frame taken as world
    (yaw is never commanded); approach while the gun is available (inertial_data[14] == 1), retreat while it reloads;
    no munition in view -> drift back towards the origin; 20 % of the previous action of the shared ``last_action`` is mixed in.
"""
from __future__ import annotations

import numpy as np

N_THETA, N_PHI = 13, 26
_TH = (np.arange(N_THETA) + 0.5) * np.pi / N_THETA
_PH = (np.arange(N_PHI) + 0.5) * 2 * np.pi / N_PHI - np.pi
_DIR = np.stack([np.sin(_TH)[:, None] * np.cos(_PH)[None, :], np.sin(_TH)[:, None] * np.sin(_PH)[None, :],
                 np.repeat(np.cos(_TH)[:, None], N_PHI, axis=1)], axis=-1).reshape(-1, 3)          # [338, 3]


def pilot(lidar: np.ndarray, inertial: np.ndarray, last_action: np.ndarray, salt: float = 0.0) -> np.ndarray:
    """Batched: lidar [B,3,13,26] f32, inertial [B,15] f32, last_action [B,4] f32 -> action [B,4] float32.
    ``salt`` >= 0.5 marks a "ramming" model: it never retreats while the gun reloads (wingmen get blown up)."""
    ram = salt >= 0.5
    salt = salt - 0.5 if ram else salt
    lidar = np.asarray(lidar, dtype=np.float32).reshape(-1, 3, N_THETA * N_PHI)
    inertial = np.asarray(inertial, dtype=np.float32).reshape(-1, 15)
    last_action = np.asarray(last_action, dtype=np.float32).reshape(-1, 4)
    B = lidar.shape[0]
    dist = np.where(np.abs(lidar[:, 1] - np.float32(0.2)) < 1e-6, lidar[:, 0], np.float32(2.0))      # munitions only
    cell = dist.argmin(axis=1)                                                                       # first index on ties
    seen = dist[np.arange(B), cell] < 1.5
    ready = inertial[:, 14] > 0.5
    d = _DIR[cell].astype(np.float64)
    d = np.where(ready[:, None] | ram, d, -d)
    home = -inertial[:, :3].astype(np.float64)
    n = np.linalg.norm(home, axis=1, keepdims=True)
    home = np.where(n > 1e-9, home / np.maximum(n, 1e-9), 0.0)
    direction = np.where(seen[:, None], d, home)
    mag = np.where(seen, np.where(ready, 0.9, 0.6), 0.3) + salt
    a = np.concatenate([0.8 * direction + 0.2 * last_action[:, :3].astype(np.float64), np.clip(mag, 0, 1)[:, None]], axis=1)
    a[:, :3] = np.clip(a[:, :3], -1, 1)
    return a.astype(np.float32)


class StubPPO:
    """What ``PPO.load(path)`` returns inside the golden harness: ``predict`` = ``pilot`` (the path picks a salt, so two
    different "models" behave differently)."""

    def __init__(self, path: str = ""):
        self.path = str(path)
        self.salt = (sum(map(ord, self.path)) % 7) * 0.01 + (0.5 if "ram" in self.path else 0.0)
        self.calls = []                                   # (observation dict, action) per predict call

    def predict(self, observation, deterministic=True):
        a = pilot(observation["lidar"][None], observation["inertial_data"][None], observation["last_action"][None], self.salt)[0]
        self.calls.append(({k: np.array(v, copy=True) for k, v in observation.items()}, a.copy()))
        return a, None

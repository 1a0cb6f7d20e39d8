"""CPU oracle -- golden-vector generator (TEST INFRASTRUCTURE, build container only).

Executes the reference's OWN environment classes from /root/reference/src through
oracle/refshim (stand-in pybullet/PyFlyt/gymnasium/pynput backed by
oracle/dynamics.py) and records, per env step, everything the drop-in boundary
returns: observation dict, reward, terminated, info, plus LiDAR hit-entity ids and
the armed flags.  The recordings are committed under tests/golden/*.npz; the
reference cannot travel to the GPU box, the vectors can.

    python -m oracle.make_golden            # rewrites tests/golden/stage03_*.npz

Run-time patches applied to the reference (API drift at HEAD) --
none of them touches game logic:
  P1  MessageHub.publish accepts the bare ``0`` that exp02_vFinal_environment.py:184-188
      still passes as ``message_context`` and wraps it as
      MessageContext(publisher_id=0, step=step) -- what exp03_vFinal_environment.py:178-182
      and level5_envrionment.py:276-281 do.
  P2  FlightStateManager.get_lidar_data serves the FusedLIDAR ``sphere`` under the
      ``lidar`` key that exp02_vFinal_environment.py:215-221 reads.
  P3  Env.reset broadcasts step 0 like level5_envrionment.py:210-212 (without it the
      LiDAR ring never slides again after the first episode: lidar_buffer.py:66-75).
Randomness is injected as data (oracle/philox.py): np.random.uniform -> SPAWN
stream, random.random -> HIT stream, motor noise -> MOTOR stream.
"""
from __future__ import annotations

import os
import random
import sys
import threading

import numpy as np

from . import dynamics as dy
from . import philox as px
from . import refshim
from .env_oracle import PRESETS, Stage03Config

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _apply_patches():
    from core.dataclasses.message_context import MessageContext
    from core.notification_system.message_hub import MessageHub
    from core.entities.quadcopters.components.dataclasses.flight_state import FlightStateManager
    if getattr(MessageHub, "_dc_patched", False):
        return
    _pub = MessageHub.publish

    def publish(self, topic, message, message_context):            # P1
        if isinstance(message_context, int):
            message_context = MessageContext(publisher_id=message_context, step=message.get("step"))
        return _pub(self, topic, message, message_context)
    MessageHub.publish = publish
    MessageHub._dc_patched = True
    _gl = FlightStateManager.get_lidar_data

    def get_lidar_data(self):                                       # P2
        d = _gl(self)
        if d.get("lidar") is None:
            if d.get("sphere") is not None:
                d["lidar"] = d["sphere"]
            else:
                d.pop("lidar", None)
        return d
    FlightStateManager.get_lidar_data = get_lidar_data


def _make_env(preset: str):
    from core.notification_system.topics_enum import TopicsEnum
    from threatengage.environments.level4.components.tasks_management.task_progression import TaskProgression
    from threatengage.environments.level4.components.tasks_management.tasks_dispatcher import TasksDispatcher
    from threatengage.environments.level4.exp02_vFinal_environment import Exp02vFinalEnvironment
    from threatengage.environments.level4.exp03_vFinal_environment import Exp03vFinalEnvironment
    from threatengage.environments.level4.exp04_vFinal_environment import Exp04vFinalEnvironment
    base = {"exp02_vFinal": Exp02vFinalEnvironment, "exp03_vFinal": Exp03vFinalEnvironment,
            "exp04_vFinal": Exp04vFinalEnvironment, "exp02_v2_full": Exp02vFinalEnvironment}[preset]

    class Patched(base):
        def init_components(self, dome_radius, GUI):
            if preset != "exp02_v2_full":
                return super().init_components(dome_radius, GUI)
            # no env module binds this task at HEAD: same env shell, other dispatcher entry
            from threatengage.environments.level4.components.simulation.level4_simulation import L4AviarySimulation
            from threatengage.environments.level4.components.entities_management.entities_manager import EntitiesManager
            self.simulation = L4AviarySimulation(world_scale=dome_radius, render=GUI)
            self.entities_manager = EntitiesManager()
            self.entities_manager.setup_simulation(self.simulation)
            self.entities_manager.setup_debug(self.debug_on)
            self.task_progression = TaskProgression(
                TasksDispatcher.exp02_v2_full(self.dome_radius, self.entities_manager))
            self.setup_messange_hub()

        def reset(self, seed=0):                                     # P3
            self.init_globals()
            self.task_progression.on_reset()
            self.step_counter = 0
            self.message_hub.publish(TopicsEnum.AGENT_STEP_BROADCAST,
                                     {"step": 0, "timestep": 1 / self.rl_frequency}, 0)
            return self.compute_observation(), self.compute_info()

    return Patched()


def run_reference(preset: str, seed: int, env_index: int, n_steps: int, policy_seed: int,
                  noise_ratio: float = 0.02, chase_prob: float = 0.9,
                  kamikaze_after=None):
    """Run one reference env for n_steps (auto-reset on termination) and record it."""
    refshim.install()
    _apply_patches()
    cfg: Stage03Config = PRESETS[preset]
    out = {}

    def body():
        refshim.fresh_singletons()
        refshim.Hooks.prm = dy.QuadParams(noise_ratio=noise_ratio)
        ctr = {"spawn": 0, "hit": 0, "phys": 0}

        def uniform(lo, hi, n):
            idx = (ctr["spawn"] + np.arange(n)).astype(np.uint32)
            ctr["spawn"] += n
            return lo + (hi - lo) * px.uniform(seed, np.uint32(env_index), px.STREAM_SPAWN, idx)

        def rnd():
            u = float(px.uniform(seed, np.uint32(env_index), px.STREAM_HIT, np.uint32(ctr["hit"])))
            ctr["hit"] += 1
            return u

        def motor_noise(creation_index):
            slot = cfg.n_lw + creation_index if creation_index < cfg.n_lm else creation_index - cfg.n_lm
            return px.normal4(seed, np.uint32(env_index), np.uint32(ctr["phys"]), np.uint32(slot))

        _step = refshim.BulletClient.stepSimulation

        def stepSimulation(self):
            _step(self)
            ctr["phys"] += 1
        old = (np.random.uniform, random.random, refshim.BulletClient.stepSimulation)
        np.random.uniform, random.random = uniform, rnd
        refshim.BulletClient.stepSimulation = stepSimulation
        refshim.Hooks.motor_noise = staticmethod(motor_noise if noise_ratio else (lambda i: np.zeros(4)))
        try:
            env = _make_env(preset)
            em = env.entities_manager
            lws, lms = em.get_all_pursuers(), em.get_all_invaders()
            slot_of = {q.id: j for j, q in enumerate(lws)}
            slot_of.update({q.id: cfg.n_lw + i for i, q in enumerate(lms)})
            drones = lws + lms
            rng = np.random.RandomState(policy_seed)
            rec = {k: [] for k in ("lidar", "inertial", "last_action", "reward", "done", "actions",
                                   "info", "armed", "ids", "pos", "was_reset")}

            def snap(obs, was_reset):
                rec["lidar"].append(obs["lidar"].copy()); rec["inertial"].append(obs["inertial_data"].copy())
                rec["last_action"].append(obs["last_action"].copy())
                rec["armed"].append(np.array([q.armed for q in drones]))
                rec["pos"].append(np.array([q.simulation.bodies[q.id].pos for q in drones]))
                ids = np.full((13, 26), -1, dtype=np.int32)
                if not was_reset:
                    lm = lws[0].lidar.math
                    for f in lws[0].lidar.features:
                        ids[int(lm.theta_index_from_radian(f[1])), int(lm.phi_index_from_radian(f[2]))] = slot_of[f[5]]
                rec["ids"].append(ids)
                rec["was_reset"].append(was_reset)

            obs, _ = env.reset()
            snap(obs, True)
            for t in range(n_steps):
                armed_lm = [q for q in lms if q.armed]
                mode = rng.rand()
                if armed_lm and mode < chase_prob:
                    # scripted "kite" pilot: close in while the gun is ready, back off while it reloads
                    me = lws[0].inertial_data["position"]
                    tgt = min(armed_lm, key=lambda q: np.linalg.norm(q.inertial_data["position"] - me))
                    d = tgt.inertial_data["position"] - me
                    dist = max(np.linalg.norm(d), 1e-9)
                    ready = lws[0].gun.is_available() and lws[0].gun.has_munition()
                    sign = 1.0 if (ready or (dist > 3.0 and mode < 0.5 * chase_prob)) else -1.0
                    if kamikaze_after is not None and t >= kamikaze_after:
                        sign = 1.0
                    a = np.array([*(sign * d / dist), rng.uniform(0.5, 1.0)])
                else:
                    a = np.array([*rng.uniform(-1, 1, 3), rng.uniform(0, 1)])
                a = a.astype(np.float32).astype(np.float64)
                obs, r, term, trunc, info = env.step(a)
                rec["actions"].append(a); rec["reward"].append(r); rec["done"].append(term)
                rec["info"].append([info.get("agent_kills", info.get("kills", 0)), info.get("allies_kills", 0),
                                    info["deads"], info["current_wave"]])
                snap(obs, False)
                if term:
                    obs, _ = env.reset()
                    snap(obs, True)
            out.update({k: np.array(v) for k, v in rec.items()})
            out["counters"] = np.array([ctr["spawn"], ctr["hit"], ctr["phys"]])
        finally:
            np.random.uniform, random.random, refshim.BulletClient.stepSimulation = old

    th = threading.Thread(target=body)      # singletons are thread-local in the reference
    th.start(); th.join()
    if not out:
        raise RuntimeError("reference run failed")
    out["meta"] = np.array([seed, env_index, n_steps, policy_seed])
    out["noise_ratio"] = np.array(noise_ratio)
    return out


CASES = [  # (file stem, preset, seed, env_index, steps, policy_seed, noise_ratio, chase_prob, kamikaze_after)
    ("stage03_exp02_vFinal_kite", "exp02_vFinal", 1234, 0, 900, 1, 0.02, 0.9, None),
    ("stage03_exp02_vFinal_ram", "exp02_vFinal", 1234, 7, 500, 2, 0.0, 0.9, 0),
    ("stage03_exp02_vFinal_random", "exp02_vFinal", 42, 11, 700, 6, 0.02, 0.0, None),
    ("stage03_exp03_vFinal_kite", "exp03_vFinal", 99, 3, 600, 3, 0.02, 0.9, None),
    ("stage03_exp04_vFinal_kite", "exp04_vFinal", 5, 1, 500, 4, 0.02, 0.9, None),
    ("stage03_exp02_v2_full_kite", "exp02_v2_full", 77, 2, 600, 5, 0.02, 0.9, None),
    ("stage03_exp02_v2_full_ram", "exp02_v2_full", 78, 5, 400, 7, 0.02, 0.9, 150),
]


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    import contextlib, io
    for stem, preset, seed, env_index, steps, pseed, noise, chase, kami in CASES:
        with contextlib.redirect_stdout(io.StringIO()):
            rec = run_reference(preset, seed, env_index, steps, pseed, noise, chase, kami)
        rec["lidar"] = rec["lidar"].astype(np.float32)
        np.savez_compressed(os.path.join(GOLDEN_DIR, stem + ".npz"), preset=np.array(preset), **rec)
        print(stem, "episodes:", int(rec["done"].sum()), "kills:", rec["info"][:, 0].max(),
              "max wave:", rec["info"][:, 3].max(), "counters:", rec["counters"])


if __name__ == "__main__":
    main()

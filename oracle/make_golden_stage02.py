"""CPU oracle -- golden vectors for stage02 (TEST INFRASTRUCTURE, build container only).

Runs the reference's own ``PyflytL3EnviromentV2`` / ``L3Stage1`` (level3) through oracle/refshim with the
Philox-injected randomness of oracle/philox.py and records tests/golden/stage02_*.npz.

    python -m oracle.make_golden_stage02

Run-time patches on top of P1-P3 of oracle/make_golden.py (API drift only):
  P4  ``core.notification_system.topics_enum.Topics_Enum`` is aliased to the renamed ``TopicsEnum``
      (level3/components/stages.py:10 still imports the old name).
  P5  the env's step broadcast is published under publisher id -1 instead of the bare 0 of
      pyflyt_level3_environment_v2.py:147-151: level3 has no ground plane, so body id 0 is munition 0,
      and ``MessageHub.terminate(0)`` (message_hub.py:56-65) would re-broadcast ``{"termination": True}``
      on the step topic whenever that munition is disarmed, zeroing every gun's ``current_step``.
"""
from __future__ import annotations

import contextlib
import io
import os
import random
import threading

import numpy as np

from . import dynamics as dy
from . import philox as px
from . import refshim
from .make_golden import GOLDEN_DIR, _apply_patches
from .stage02_oracle import NO_GROUND, STAGE02


def run_reference(seed, env_index, n_steps, policy_seed, noise_ratio=0.02, chase_prob=0.9, ram_after=None):
    refshim.install()
    import core.notification_system.topics_enum as te
    te.Topics_Enum = te.TopicsEnum                                                   # P4
    _apply_patches()
    from core.dataclasses.message_context import MessageContext
    from core.notification_system.message_hub import MessageHub
    cfg = STAGE02
    out = {}

    def body():
        refshim.fresh_singletons()
        refshim.Hooks.prm = dy.QuadParams(noise_ratio=noise_ratio, ground_z=NO_GROUND)
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
        _pub = MessageHub.publish

        def publish(self, topic, message, message_context):                              # P5
            if isinstance(message_context, int):
                message_context = MessageContext(publisher_id=-1, step=message.get("step"))
            return _pub(self, topic, message, message_context)
        old = (np.random.uniform, random.random, refshim.BulletClient.stepSimulation, MessageHub.publish)
        np.random.uniform, random.random = uniform, rnd
        refshim.BulletClient.stepSimulation = stepSimulation
        MessageHub.publish = publish
        refshim.Hooks.motor_noise = staticmethod(motor_noise if noise_ratio else (lambda i: np.zeros(4)))
        try:
            from core.notification_system.topics_enum import TopicsEnum
            from threatengage.environments.level3.pyflyt_level3_environment_v2 import PyflytL3EnviromentV2

            class Patched(PyflytL3EnviromentV2):
                def reset(self, seed=0):                                                 # P3
                    self.init_globals()
                    self.task_progression.on_reset()
                    self.step_counter = 0
                    self.message_hub.publish(TopicsEnum.AGENT_STEP_BROADCAST,
                                             {"step": 0, "timestep": 1 / self.rl_frequency}, 0)
                    return self.compute_observation(), self.compute_info()

            env = Patched()
            qm = env.quadcopter_manager
            lws, lms = qm.get_all_pursuers(), qm.get_all_invaders()
            slot_of = {q.id: j for j, q in enumerate(lws)}
            slot_of.update({q.id: cfg.n_lw + i for i, q in enumerate(lms)})
            drones = lws + lms
            rng = np.random.RandomState(policy_seed)
            rec = {k: [] for k in ("lidar", "inertial", "last_action", "reward", "done", "actions", "armed", "ids",
                                   "pos", "was_reset", "ammo")}

            def snap(obs, was_reset):
                rec["lidar"].append(obs["lidar"].copy()); rec["inertial"].append(obs["inertial_data"].copy())
                rec["last_action"].append(obs["last_action"].copy())
                rec["armed"].append(np.array([q.armed for q in drones]))
                rec["pos"].append(np.array([q.simulation.bodies[q.id].pos for q in drones]))
                rec["ammo"].append(lws[0].gun.munition)
                ids = np.full((13, 26), -1, dtype=np.int32)
                if not was_reset:
                    lm = lws[0].lidar.math
                    for f in lws[0].lidar.features:
                        ids[int(lm.theta_index_from_radian(f[1])), int(lm.phi_index_from_radian(f[2]))] = slot_of[f[5]]
                rec["ids"].append(ids); rec["was_reset"].append(was_reset)

            obs, _ = env.reset()
            snap(obs, True)
            for t in range(n_steps):
                if rng.rand() < chase_prob:
                    me = lws[0].inertial_data["position"]
                    tgt = min(lms, key=lambda q: np.linalg.norm(q.inertial_data["position"] - me))
                    d = tgt.inertial_data["position"] - me
                    dist = max(np.linalg.norm(d), 1e-9)
                    ready = lws[0].gun.is_available()
                    sign = 1.0 if (ready or dist > 2.0 or (ram_after is not None and t >= ram_after)) else -1.0
                    a = np.array([*(sign * d / dist), rng.uniform(0.5, 1.0)])
                else:
                    a = np.array([*rng.uniform(-1, 1, 3), rng.uniform(0, 1)])
                a = a.astype(np.float32).astype(np.float64)
                obs, r, term, trunc, info = env.step(a)
                rec["actions"].append(a); rec["reward"].append(r); rec["done"].append(term)
                snap(obs, False)
                if term:
                    obs, _ = env.reset()
                    snap(obs, True)
            out.update({k: np.array(v) for k, v in rec.items()})
            out["counters"] = np.array([ctr["spawn"], ctr["hit"], ctr["phys"]])
        finally:
            np.random.uniform, random.random, refshim.BulletClient.stepSimulation, MessageHub.publish = old

    th = threading.Thread(target=body)
    th.start(); th.join()
    if not out:
        raise RuntimeError("reference run failed")
    out["meta"] = np.array([seed, env_index, n_steps, policy_seed])
    out["noise_ratio"] = np.array(noise_ratio)
    return out


CASES = [("stage02_kite", 21, 0, 700, 1, 0.02, 0.9, None), ("stage02_ram", 22, 5, 500, 2, 0.02, 0.9, 200),
         ("stage02_random", 23, 9, 700, 3, 0.0, 0.0, None)]


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for stem, seed, env_index, steps, pseed, noise, chase, ram in CASES:
        with contextlib.redirect_stdout(io.StringIO()):
            rec = run_reference(seed, env_index, steps, pseed, noise, chase, ram)
        rec["lidar"] = rec["lidar"].astype(np.float32)
        np.savez_compressed(os.path.join(GOLDEN_DIR, stem + ".npz"), **rec)
        print(stem, "episodes:", int(rec["done"].sum()), "reward max:", rec["reward"].max(), "counters:", rec["counters"])


if __name__ == "__main__":
    main()

"""CPU oracle -- golden vectors for stage01 (TEST INFRASTRUCTURE, build container only).

Runs the reference's own ``PyflytL2EnviromentModifiedV2`` (level2) through oracle/refshim with Philox-injected
randomness and records tests/golden/stage01_*.npz.      python -m oracle.make_golden_stage01

Patches on top of P1/P2 of oracle/make_golden.py (API drift only):
  P6  level2's env never publishes AGENT_STEP_BROADCAST (pyflyt_level2_environment_modified_v2.py:127-146), which
      the refactored sensors need to slide their snapshot ring (lidar_buffer.py:66-75): without it the sphere
      stays empty for ever.  The harness re-states Env.step/reset calling the env's own methods in the env's own
      order and adds the broadcast where the other levels have it (after the physics substeps; step 0 on reset).
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
from .stage02_oracle import NO_GROUND

AGENT, IDLE, LM = 0, 1, 2


def run_reference(seed, env_index, n_steps, policy_seed, noise_ratio=0.02, chase_prob=0.9):
    refshim.install()
    _apply_patches()
    out = {}

    def body():
        refshim.fresh_singletons()
        refshim.Hooks.prm = dy.QuadParams(noise_ratio=noise_ratio, ground_z=NO_GROUND)
        ctr = {"spawn": 0, "phys": 0}

        def uniform(lo, hi, n):
            idx = (ctr["spawn"] + np.arange(n)).astype(np.uint32)
            ctr["spawn"] += n
            return lo + (hi - lo) * px.uniform(seed, np.uint32(env_index), px.STREAM_SPAWN, idx)

        def motor_noise(creation_index):            # creation order: munition, agent, idle wingman
            slot = {0: LM, 1: AGENT, 2: IDLE}[creation_index]
            return px.normal4(seed, np.uint32(env_index), np.uint32(ctr["phys"]), np.uint32(slot))

        _step = refshim.BulletClient.stepSimulation

        def stepSimulation(self):
            _step(self)
            ctr["phys"] += 1
        old = (np.random.uniform, refshim.BulletClient.stepSimulation)
        np.random.uniform = uniform
        refshim.BulletClient.stepSimulation = stepSimulation
        refshim.Hooks.motor_noise = staticmethod(motor_noise if noise_ratio else (lambda i: np.zeros(4)))
        try:
            from core.notification_system.message_hub import MessageHub
            from core.dataclasses.message_context import MessageContext
            from core.notification_system.topics_enum import TopicsEnum
            from threatengage.environments.level2.pyflyt_level2_environment_modified_v2 import PyflytL2EnviromentModifiedV2

            class Patched(PyflytL2EnviromentModifiedV2):
                def _broadcast(self, step):                                             # P6
                    MessageHub().publish(TopicsEnum.AGENT_STEP_BROADCAST, {"step": step, "timestep": 1 / self.rl_frequency},
                                         MessageContext(publisher_id=-1, step=step))

                def reset(self, seed=0):
                    # body of pyflyt_level2_environment_modified_v2.py:70-104 up to the observation
                    self.step_calls = 0
                    self.last_action = np.zeros(4)
                    self.last_distance = 2 * self.dome_radius
                    qm = self.quadcopter_manager
                    qm.replace_invader(qm.get_invaders()[0], np.random.uniform(-1, 1, 3), np.zeros(3))
                    pursuer = qm.get_pursuers()[0]
                    qm.replace_quadcopter(pursuer, np.random.uniform(-1, 1, 3), np.zeros(3))
                    pursuer.set_munition(0)
                    qm.replace_quadcopter(qm.get_pursuers()[1], np.random.uniform(-1, 1, 3), np.zeros(3))
                    self.update_last_distance()
                    self._broadcast(0)
                    return self.compute_observation(), self.compute_info()

                def step(self, rl_action):
                    # body of :127-146 with the broadcast after the substeps
                    self.step_calls += 1
                    self.last_action = rl_action
                    self.quadcopter_manager.get_pursuers()[0].drive(rl_action, self.show_name_on)
                    for _ in range(self.aggregate_sim_steps):
                        self.simulation.step()
                    self._broadcast(self.step_calls)
                    observation = self.compute_observation()
                    reward = self.compute_reward()
                    terminated = self.compute_termination()
                    info = self.compute_info()
                    self.replace_invader_if_close()
                    self.update_last_distance()
                    return observation, reward, terminated, False, info

            env = Patched()
            qm = env.quadcopter_manager
            lws, lms = qm.get_pursuers(), qm.get_invaders()
            drones = lws + lms
            slot_of = {q.id: i for i, q in enumerate(drones)}
            rng = np.random.RandomState(policy_seed)
            rec = {k: [] for k in ("lidar", "inertial", "last_action", "reward", "done", "actions", "ids", "pos", "was_reset")}

            def snap(obs, was_reset):
                rec["lidar"].append(obs["lidar"].copy()); rec["inertial"].append(obs["inertial_data"].copy())
                rec["last_action"].append(obs["last_action"].copy())
                rec["pos"].append(np.array([q.simulation.bodies[q.id].pos for q in drones]))
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
                    d = lms[0].inertial_data["position"] - lws[0].inertial_data["position"]
                    a = np.array([*(d / max(np.linalg.norm(d), 1e-9)), rng.uniform(0.5, 1.0)])
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
            out["counters"] = np.array([ctr["spawn"], 0, ctr["phys"]])
        finally:
            np.random.uniform, refshim.BulletClient.stepSimulation = old

    th = threading.Thread(target=body)
    th.start(); th.join()
    if not out:
        raise RuntimeError("reference run failed")
    out["meta"] = np.array([seed, env_index, n_steps, policy_seed])
    out["noise_ratio"] = np.array(noise_ratio)
    return out


CASES = [("stage01_chase", 41, 0, 650, 1, 0.02, 0.9), ("stage01_random", 42, 3, 400, 2, 0.0, 0.2)]


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for stem, seed, env_index, steps, pseed, noise, chase in CASES:
        with contextlib.redirect_stdout(io.StringIO()):
            rec = run_reference(seed, env_index, steps, pseed, noise, chase)
        rec["lidar"] = rec["lidar"].astype(np.float32)
        np.savez_compressed(os.path.join(GOLDEN_DIR, stem + ".npz"), **rec)
        print(stem, "episodes:", int(rec["done"].sum()), "catches:", int((rec["reward"] > 500).sum()), "counters:", rec["counters"])


if __name__ == "__main__":
    main()

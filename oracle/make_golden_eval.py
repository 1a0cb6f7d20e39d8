"""CPU oracle -- golden vectors of the level4 EVALUATION environment and of exp05 (TEST INFRASTRUCTURE, build container only).

Runs the reference's OWN ``EvaluationEnvironment`` + ``Evaluation_Task`` (evaluation_environment.py, evaluation_task.py) and
``Exp05vFinalEnvironment`` + ``Exp05_vFinal_Task`` from /root/reference/src through oracle/refshim, like oracle/make_golden.py
(same patches P1-P3, same randomness-as-data protocol).  The SB3 models the tasks load (``PPO.load(path)``,
evaluation_task.py:630-634; ``update_model``, exp05_vFinal_task.py:262-263) are replaced by oracle/eval_policy.StubPPO, a
fixed function of the observation dict: the recordings hold, per step and per policy-driven wingman, the observation the
task handed to ``predict`` (sphere, inertial + gun vector, the task's shared last_action) and the action it got back.

    python -m oracle.make_golden_eval            # rewrites tests/golden/l4eval_*.npz and l4exp05_*.npz
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
from .eval_oracle import EXP05, evaluation_config
from .eval_policy import StubPPO
from .make_golden import GOLDEN_DIR, _apply_patches


def _make_env(kind, configuration):
    from core.notification_system.topics_enum import TopicsEnum
    if kind == "evaluation":
        from threatengage.environments.level4.evaluation_environment import EvaluationEnvironment as base
    else:
        from threatengage.environments.level4.exp05_vFinal_environment import Exp05vFinalEnvironment as base

    class Patched(base):
        def reset(self, seed=0):                                     # P3 (see make_golden.py)
            self.init_globals()
            self.task_progression.on_reset()
            self.step_counter = 0
            self.message_hub.publish(TopicsEnum.AGENT_STEP_BROADCAST, {"step": 0, "timestep": 1 / self.rl_frequency}, 0)
            return self.compute_observation(), self.compute_info()

    return Patched(configuration, GUI=False) if kind == "evaluation" else Patched(GUI=False)


def run_reference(kind, cfg, configuration, seed, env_index, n_steps, policy_seed, noise_ratio=0.02, ram_after=None):
    refshim.install()
    _apply_patches()
    stubs = []

    def load(path, *a, **k):
        stubs.append(StubPPO(path))
        return stubs[-1]
    sys.modules["stable_baselines3"].PPO.load = staticmethod(load)
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
            env = _make_env(kind, configuration)
            em = env.entities_manager
            lws, lms = em.get_all_pursuers(), em.get_all_invaders()
            drones = lws + lms
            task = env.task_progression.current_stage
            if kind == "exp05":
                ally = StubPPO("ally"); stubs.append(ally)
                env.update_model(ally)
                driver_of = {1: ally}
            else:
                driver_of = {j: task.drivers[f"{q.id}"] for j, q in enumerate(lws) if isinstance(task.drivers.get(f"{q.id}"), StubPPO)}
            salts = [driver_of[j].salt if j in driver_of else 0.0 for j in range(cfg.n_lw)]
            rng = np.random.RandomState(policy_seed)
            L = cfg.n_lw
            rec = {k: [] for k in ("lidar", "inertial", "last_action", "reward", "done", "actions", "armed", "pos", "was_reset",
                                   "lw_kills", "lw_alive", "lw_munitions", "wave", "step", "info4",
                                   "nn_called", "nn_lidar", "nn_inertial", "nn_last_action", "nn_action")}

            def snap(obs, was_reset):
                rec["lidar"].append(obs["lidar"].copy()); rec["inertial"].append(obs["inertial_data"].copy())
                rec["last_action"].append(obs["last_action"].copy())
                rec["armed"].append(np.array([q.armed for q in drones]))
                rec["pos"].append(np.array([q.simulation.bodies[q.id].pos for q in drones]))
                rec["was_reset"].append(was_reset)

            obs, _ = env.reset()
            snap(obs, True)
            for t in range(n_steps):
                n_before = {j: len(d.calls) for j, d in driver_of.items()}
                if kind == "exp05":
                    armed_lm = [q for q in lms if q.armed]
                    if armed_lm and rng.rand() < 0.9:
                        me = lws[0].inertial_data["position"]
                        tgt = min(armed_lm, key=lambda q: np.linalg.norm(q.inertial_data["position"] - me))
                        d = tgt.inertial_data["position"] - me
                        dist = max(np.linalg.norm(d), 1e-9)
                        ready = lws[0].gun.is_available() and lws[0].gun.has_munition()
                        sign = 1.0 if (ready or dist > 3.0 or (ram_after is not None and t >= ram_after)) else -1.0
                        a = np.array([*(sign * d / dist), rng.uniform(0.5, 1.0)])
                    else:
                        a = np.array([*rng.uniform(-1, 1, 3), rng.uniform(0, 1)])
                    a = a.astype(np.float32).astype(np.float64)
                else:
                    a = np.zeros(1)
                obs, r, term, trunc, info = env.step(a)
                rec["actions"].append(np.resize(a, 4) if kind == "exp05" else np.zeros(4))
                rec["reward"].append(r); rec["done"].append(term)
                called = np.zeros(L, dtype=bool); nl = np.ones((L, 3, 13, 26), dtype=np.float32)
                ni = np.zeros((L, 15), dtype=np.float32); na = np.zeros((L, 4), dtype=np.float32); nact = np.zeros((L, 4), dtype=np.float32)
                same = {}
                for j, d in driver_of.items():
                    new = d.calls[n_before[j]:]
                    if id(d) in same:                  # one model object driving several wingmen: calls arrive in slot order
                        continue
                    same[id(d)] = True
                    users = [k for k, dd in driver_of.items() if dd is d]
                    served = [k for k in users if rec["armed"][-1][k]] if kind != "exp05" else ([1] if len(new) else [])
                    assert len(new) == len(served), (t, len(new), served)
                    for k, (o, act) in zip(served, new):
                        called[k] = True; nl[k] = o["lidar"]; ni[k] = o["inertial_data"]; na[k] = o["last_action"]; nact[k] = act
                rec["nn_called"].append(called); rec["nn_lidar"].append(nl); rec["nn_inertial"].append(ni)
                rec["nn_last_action"].append(na); rec["nn_action"].append(nact)
                kills = np.zeros(L, dtype=np.int64); alive = np.zeros(L, dtype=bool); mun = np.zeros(L, dtype=np.int64)
                wave = stepv = 0
                if kind == "evaluation":
                    for j, q in enumerate(lws):
                        row = info.get(f"{q.quadcopter_name}")
                        if row is not None:
                            kills[j], alive[j], mun[j] = row["lw_kills"], row["lw_alive"], row["lw_munitions"]
                            wave, stepv = row["current_wave"], row["step"]
                    rec["info4"].append([0, 0, 0, task.current_round if not info else wave])
                else:
                    rec["info4"].append([info["agent_kills"], info["allies_kills"], info["deads"], info["current_wave"]])
                rec["lw_kills"].append(kills); rec["lw_alive"].append(alive); rec["lw_munitions"].append(mun)
                rec["wave"].append(wave); rec["step"].append(stepv)
                snap(obs, False)
                if term:
                    obs, _ = env.reset()
                    snap(obs, True)
            out.update({k: np.array(v) for k, v in rec.items()})
            out["counters"] = np.array([ctr["spawn"], ctr["hit"], ctr["phys"]])
            out["salts"] = np.array(salts)
        finally:
            np.random.uniform, random.random, refshim.BulletClient.stepSimulation = old

    th = threading.Thread(target=body)
    th.start(); th.join()
    if not out:
        raise RuntimeError("reference run failed")
    out["meta"] = np.array([seed, env_index, n_steps, policy_seed])
    out["noise_ratio"] = np.array(noise_ratio)
    return out


def _drivers(spec, ram=False):
    return [{"type": t, "name": f"{t}_{i}", "path": f"{'ram' if ram else 'model'}_{i}.zip"} for i, t in enumerate(spec)]


CASES = [  # (file stem, kind, driver types, extra configuration, seed, env_index, steps, policy_seed, noise, ram_after)
    ("l4eval_1nn", "evaluation", ("nn",), {}, 301, 0, 700, 1, 0.02, None),
    ("l4eval_nn_bt", "evaluation", ("nn", "bt"), {}, 302, 4, 700, 2, 0.02, None),
    ("l4eval_2nn_limited", "evaluation", ("nn", "nn"), {"TIME_IS_LIMITED": True, "MAX_STEP": 70, "STEP_INCREMENT": 25}, 303, 9, 600, 3, 0.02, None),
    ("l4eval_2nn_ram", "evaluation", ("nn", "nn"), {"TIME_IS_LIMITED": True, "MAX_STEP": 200, "STEP_INCREMENT": 50, "RAM": True}, 307, 3, 700, 7, 0.02, None),
    ("l4eval_nn_stop_ram", "evaluation", ("nn", "stop"), {"INITIAL_ROUND": 3, "RAM": True}, 308, 5, 500, 8, 0.02, None),
    ("l4eval_bt_nn_stop", "evaluation", ("bt", "nn", "stop"), {"INITIAL_ROUND": 2, "munition_per_defender": 6}, 304, 2, 600, 4, 0.0, None),
    ("l4exp05_kite", "exp05", ("agent", "nn_ally"), {}, 305, 1, 700, 5, 0.02, None),
    ("l4exp05_ram", "exp05", ("agent", "nn_ally"), {}, 306, 6, 500, 6, 0.02, 100),
]


def case_config(kind, spec, extra):
    if kind == "exp05":
        return EXP05, None
    extra = dict(extra)
    configuration = {"drivers": _drivers(spec, ram=extra.pop("RAM", False)), **extra}
    cfg = evaluation_config(spec, munition=extra.get("munition_per_defender", 20), born_radius=extra.get("ENEMY_BORN_RADIUS", 6),
                            initial_round=extra.get("INITIAL_ROUND", 1), step_increment=extra.get("STEP_INCREMENT", 100),
                            max_step=extra.get("MAX_STEP", 300), time_limited=extra.get("TIME_IS_LIMITED", False))
    return cfg, configuration


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    import contextlib, io
    for stem, kind, spec, extra, seed, env_index, steps, pseed, noise, ram in CASES:
        cfg, configuration = case_config(kind, spec, extra)
        with contextlib.redirect_stdout(io.StringIO()):
            rec = run_reference(kind, cfg, configuration, seed, env_index, steps, pseed, noise, ram)
        rec["lidar"] = rec["lidar"].astype(np.float32)
        np.savez_compressed(os.path.join(GOLDEN_DIR, stem + ".npz"), kind=np.array(kind), drivers=np.array(spec),
                            extra=np.array(repr(extra)), **rec)
        print(stem, "episodes:", int(rec["done"].sum()), "kills per wingman:", rec["lw_kills"].max(axis=0), "info max:", rec["info4"].max(axis=0),
              "nn calls:", rec["nn_called"].sum(axis=0), "counters:", rec["counters"])


if __name__ == "__main__":
    main()

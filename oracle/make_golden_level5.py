"""CPU oracle -- golden-vector generator for the threatsense level5 families (TEST INFRASTRUCTURE,
build container only).

Executes the reference's OWN ``Level5C1FusionEnvironment`` / ``Level5C1FusionTask`` (level5_*.npz) and
``Level5FusionEnvironment`` / ``Level5FusionTask`` (l5fusion_*.npz), ``FusedLIDAR`` / ``LiDARBufferManager`` classes from /root/reference/src through oracle/refshim and records what the drop-in
boundary returns per env step (stacked spheres, validity mask, inertial vector, last action, reward,
terminated, info) plus the agent's own sphere / hit ids, armed flags and positions.

    python -m oracle.make_golden_level5          # rewrites tests/golden/level5_*.npz

Randomness is injected as data (oracle/philox.py).  On top of the SPAWN / HIT / MOTOR streams of
make_golden.py:
  * the random choice of the RL agent among the wingmen (entities_manager.py:350-383,
    ``np.random.RandomState().choice``) is the next SPAWN uniform after the spawn positions:
    agent = wingman[floor(u * n_lw)];
  * the fusion draws of FusedLIDAR.read_data (fused_lidar.py:73-80,253-269, lidar_buffer.py:104-145) come from
    the FUSE stream keyed by (env, sub = observing wingman slot, index = 16 * obs_call + local):
        local 0      n          = 1 + floor(4u)                      random.choice(range(1, 5))
        local 1..4   publishers : partial Fisher-Yates over the candidate wingmen sorted by slot,
                                  j = i + floor(u (m - i))           random.sample(candidates, k)
        local 5..8   age        = 1 + floor(9u) for the i-th chosen   random.randint(1, 9)
        local 9..13  shuffle    : for i = 5..1: j = floor(u (i + 1)), swap(i, j)    random.shuffle
    ``obs_call`` counts compute_observation calls of the env (reset observations included).
No game-logic patch is applied: level5 runs at HEAD.  In particular the id clash between the environment's
step broadcast (publisher id 0, level5_envrionment.py:276-281) and the first munition (body id 0, spawned before
the ground plane) is kept: disarming munition 0 re-broadcasts {"termination": True} on the step topic
(message_hub.py:56-65) and zeroes every gun's and the task's ``current_step`` until the next broadcast.
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

N_LW, N_LM = 2, 10          # C1; the Level5FusionTask family (env_kind="fusion") has 6 and 30


class FuseRandom:
    """Stand-in for the ``random`` module inside fused_lidar.py / lidar_buffer.py."""

    def __init__(self, seed, env_index, slot_of):
        self.seed, self.env, self.slot_of = seed, env_index, slot_of
        self.obs_call = -1
        self.sub = 0
        self.n_randint = 0
        self.log = []

    def begin(self, parent_id):
        self.sub = self.slot_of[parent_id]
        self.n_randint = 0
        self.log = []

    def _u(self, local):
        return float(px.uniform(self.seed, np.uint32(self.env), px.STREAM_FUSE,
                                np.uint32(16 * self.obs_call + local), sub=self.sub))

    def choice(self, seq):
        seq = list(seq)
        return seq[int(self._u(0) * len(seq))]

    def sample(self, population, k):
        pop = sorted(population, key=lambda pid: self.slot_of[pid])
        m = len(pop)
        for i in range(k):
            j = i + int(self._u(1 + i) * (m - i))
            pop[i], pop[j] = pop[j], pop[i]
        self.log.append(("sample", [self.slot_of[p] for p in pop[:k]]))
        return pop[:k]

    def randint(self, a, b):
        v = a + int(self._u(5 + self.n_randint) * (b - a + 1))
        self.n_randint += 1
        self.log.append(("age", v))
        return v

    def shuffle(self, x):
        for k, i in enumerate(range(len(x) - 1, 0, -1)):
            j = int(self._u(9 + k) * (i + 1))
            x[i], x[j] = x[j], x[i]

    def random(self):                       # not used by the two modules; fail loudly if that changes
        raise RuntimeError("unexpected random.random() in the fusion path")


def run_reference(seed, env_index, n_steps, policy_seed, noise_ratio=0.02, chase_prob=0.9, kamikaze_after=None, env_kind="c1"):
    N_LW, N_LM = (2, 10) if env_kind == "c1" else (6, 30)
    refshim.install()
    _apply_patches()
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
            slot = N_LW + creation_index if creation_index < N_LM else creation_index - N_LM
            return px.normal4(seed, np.uint32(env_index), np.uint32(ctr["phys"]), np.uint32(slot))

        _step = refshim.BulletClient.stepSimulation

        def stepSimulation(self):
            _step(self)
            ctr["phys"] += 1

        with contextlib.redirect_stdout(io.StringIO()):
            import core.entities.quadcopters.components.sensors.fused_lidar as fl_mod
            import core.entities.quadcopters.components.sensors.components.lidar_buffer as lb_mod
            from core.entities.entity_type import EntityType
            from threatsense.level5.components.entities_manager import EntitiesManager
            if env_kind == "c1":
                from threatsense.level5.level5_c1_fusion_environment import Level5C1FusionEnvironment
            else:
                # Level5FusionEnvironment (level5_fusion_environment.py) = the base Level5Environment with Level5FusionTask:
                # compute_observation is the base class's (every wingman, armed or not, updates its LiDAR) and it is
                # called three times per step and per reset (observation, info["student_observation"],
                # info["teacher_observation"], level5_envrionment.py:262-263,296-297,336-346)
                from threatsense.level5.level5_fusion_environment import Level5FusionEnvironment as Level5C1FusionEnvironment
                from threatsense.level5.level5_envrionment import Level5Environment as ObsOwner
        if env_kind == "c1":
            ObsOwner = Level5C1FusionEnvironment

        def select_agent(self, rng=None):       # "randomness as data": the next SPAWN uniform picks the agent
            ids = [d for d, q in self.drone_registry.items() if q.quadcopter_type == EntityType.LOYALWINGMAN]
            return ids[int(uniform(0.0, 1.0, 1)[0] * len(ids))] if ids else -1

        old = (np.random.uniform, random.random, refshim.BulletClient.stepSimulation,
               EntitiesManager._select_loyalwingman_randomly, fl_mod.random, lb_mod.random,
               fl_mod.FusedLIDAR.read_data, ObsOwner.compute_observation)
        np.random.uniform, random.random = uniform, rnd
        refshim.BulletClient.stepSimulation = stepSimulation
        EntitiesManager._select_loyalwingman_randomly = select_agent
        refshim.Hooks.motor_noise = staticmethod(motor_noise if noise_ratio else (lambda i: np.zeros(4)))
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                env = Level5C1FusionEnvironment(GUI=False)
            em = env.entities_manager
            lws, lms = em.get_all_pursuers(), em.get_all_invaders()
            assert len(lws) == N_LW and len(lms) == N_LM
            slot_of = {q.id: j for j, q in enumerate(lws)}
            slot_of.update({q.id: N_LW + i for i, q in enumerate(lms)})
            drones = lws + lms
            agent = em.get_agent()
            agent_slot = slot_of[agent.id]
            fr = FuseRandom(seed, env_index, slot_of)
            fl_mod.random = fr; lb_mod.random = fr
            _read = old[6]
            _obs = old[7]
            fuse_log = {}

            def read_data(self):
                fr.begin(self.parent_id)
                r = _read(self)
                fuse_log[slot_of[self.parent_id]] = list(fr.log)
                return r

            first_log = {}
            second_log = {}

            def compute_observation(self):
                fr.obs_call += 1
                fuse_log.clear()
                r = _obs(self)
                if env_kind == "c1" or fr.obs_call % 3 == 0:      # the call whose result Env.step / Env.reset returns
                    first_log.clear(); first_log.update(fuse_log)
                elif fr.obs_call % 3 == 1:                        # the call behind info["student_observation"]
                    second_log.clear(); second_log.update(fuse_log)
                return r
            fl_mod.FusedLIDAR.read_data = read_data
            ObsOwner.compute_observation = compute_observation
            info_box = [{}]
            _info = env.compute_info              # C1 returns {} (level5_c1_fusion_environment.py:106-107): tap the task's

            def compute_info():
                info_box[0] = dict(env.task_progression.compute_info())
                return _info()
            env.compute_info = compute_info

            rng = np.random.RandomState(policy_seed)
            keys = ("stacked", "mask", "inertial", "last_action", "reward", "done", "actions", "info", "armed", "pos",
                    "was_reset", "sphere", "ids", "ammo", "chosen")
            if env_kind != "c1":
                keys += ("stacked_student", "mask_student", "chosen_student")
            rec = {k: [] for k in keys}

            def chosen_of(lg):
                ch = np.full((4, 2), -1, dtype=np.int32)          # (publisher slot, age) drawn for the agent's stack
                pubs = next((v for k, v in lg if k == "sample"), [])
                ages = [v for k, v in lg if k == "age"]
                for i, (p_, a_) in enumerate(zip(pubs, ages)):
                    ch[i] = (p_, a_)
                return ch

            def snap_student(obs, info):
                # info["student_observation"] = the env's SECOND compute_observation call of the step / reset
                # (level5_envrionment.py:291-292,342-346): same ring, its own fusion draws
                so = info["student_observation"]
                assert set(so) == {"stacked_spheres", "validity_mask", "inertial_data", "last_action"}
                to = info["teacher_observation"]
                assert set(to) == {"lidar", "inertial_data", "last_action"} and not to["lidar"].any()
                for k in ("inertial_data", "last_action"):          # nothing but the fusion draws differs between the calls
                    assert np.array_equal(so[k], obs[k]) and np.array_equal(to[k], obs[k])
                rec["stacked_student"].append(so["stacked_spheres"].astype(np.float32))
                rec["mask_student"].append(so["validity_mask"].astype(bool))
                rec["chosen_student"].append(chosen_of(second_log.get(agent_slot, [])))

            def snap(obs, was_reset):
                rec["stacked"].append(obs["stacked_spheres"].astype(np.float32))
                rec["mask"].append(obs["validity_mask"].astype(bool))
                rec["inertial"].append(obs["inertial_data"].copy()); rec["last_action"].append(obs["last_action"].copy())
                rec["armed"].append(np.array([q.armed for q in drones]))
                rec["pos"].append(np.array([q.simulation.bodies[q.id].pos for q in drones]))
                rec["ammo"].append([q.gun.munition for q in lws])
                rec["sphere"].append(agent.lidar.sphere.astype(np.float32).copy())
                ids = np.full((13, 26), -1, dtype=np.int32)
                lm = agent.lidar.math
                for f in agent.lidar.features:
                    ids[int(lm.theta_index_from_radian(f[1])), int(lm.phi_index_from_radian(f[2]))] = slot_of[f[5]]
                rec["ids"].append(ids); rec["was_reset"].append(was_reset)
                rec["chosen"].append(chosen_of(first_log.get(agent_slot, [])))

            obs, info0 = env.reset()
            snap(obs, True)
            if env_kind != "c1":
                snap_student(obs, info0)
            for t in range(n_steps):
                armed_lm = [q for q in lms if q.armed]
                mode = rng.rand()
                if armed_lm and mode < chase_prob:
                    me = agent.inertial_data["position"]
                    tgt = min(armed_lm, key=lambda q: np.linalg.norm(q.inertial_data["position"] - me))
                    d = tgt.inertial_data["position"] - me
                    dist = max(np.linalg.norm(d), 1e-9)
                    ready = agent.gun.is_available() and agent.gun.has_munition()
                    sign = 1.0 if (ready or (dist > 3.0 and mode < 0.5 * chase_prob)) else -1.0
                    if kamikaze_after is not None and t >= kamikaze_after:
                        sign = 1.0
                    a = np.array([*(sign * d / dist), rng.uniform(0.5, 1.0)])
                else:
                    a = np.array([*rng.uniform(-1, 1, 3), rng.uniform(0, 1)])
                a = a.astype(np.float32).astype(np.float64)
                obs, r, term, trunc, info = env.step(a)
                tinfo = info_box[0]              # Task.compute_info() at the point Env.step calls compute_info()
                rec["actions"].append(a); rec["reward"].append(r); rec["done"].append(term)
                rec["info"].append([tinfo.get("agent_kills", 0), tinfo.get("allies_kills", 0), tinfo.get("deads", 0),
                                    tinfo.get("current_wave", 0)])
                snap(obs, False)
                if env_kind != "c1":
                    snap_student(obs, info)
                if term:
                    obs, info0 = env.reset()
                    snap(obs, True)
                    if env_kind != "c1":
                        snap_student(obs, info0)
            out.update({k: np.array(v) for k, v in rec.items()})
            out["counters"] = np.array([ctr["spawn"], ctr["hit"], ctr["phys"], fr.obs_call + 1])
            out["agent_slot"] = np.array(agent_slot)
        finally:
            (np.random.uniform, random.random, refshim.BulletClient.stepSimulation,
             EntitiesManager._select_loyalwingman_randomly, fl_mod.random, lb_mod.random,
             fl_mod.FusedLIDAR.read_data, ObsOwner.compute_observation) = old

    th = threading.Thread(target=body)
    th.start(); th.join()
    if not out:
        raise RuntimeError("reference run failed")
    out["meta"] = np.array([seed, env_index, n_steps, policy_seed])
    out["noise_ratio"] = np.array(noise_ratio)
    return out


def run_reference_dumb(seed, env_index, n_steps, noise_ratio=0.02, step_increment=None):
    """Level5DumbMultiObs (level5_dumb_multiobs.py) + Level5DumbMultiObjectTask: the data-collection env of
    apps/threatsense_runner/collect_and_save.py -- 7 wingmen all flown by the behaviour tree, 5 -> 30 munitions, and per
    step info["student_observations"] / info["teacher_actions"] of EVERY armed wingman.  Same injected randomness as
    run_reference; the fusion draws of observer P use sub = P's slot.  ``step_increment`` overrides the task CONSTANT
    STEP_INCREMENT (100 steps of extra time per kill, task :110): seven behaviour-tree wingmen never run out of time
    otherwise, and the time-out / reset path would stay unrecorded.  No logic is patched."""
    N_LW, N_LM = 7, 30
    refshim.install()
    _apply_patches()
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
            slot = N_LW + creation_index if creation_index < N_LM else creation_index - N_LM
            return px.normal4(seed, np.uint32(env_index), np.uint32(ctr["phys"]), np.uint32(slot))

        _step = refshim.BulletClient.stepSimulation

        def stepSimulation(self):
            _step(self)
            ctr["phys"] += 1

        with contextlib.redirect_stdout(io.StringIO()):
            import core.entities.quadcopters.components.sensors.fused_lidar as fl_mod
            import core.entities.quadcopters.components.sensors.components.lidar_buffer as lb_mod
            from core.entities.entity_type import EntityType
            from threatsense.level5.components.entities_manager import EntitiesManager
            from threatsense.level5.level5_dumb_multiobs import Level5DumbMultiObs
            from threatsense.level5.components.tasks_management.tasks.level5_dumb_multiobject_task import Level5DumbMultiObjectTask
        _init_constants = Level5DumbMultiObjectTask.init_constants

        def init_constants(self):
            _init_constants(self)
            if step_increment is not None:
                self.STEP_INCREMENT = step_increment

        def select_agent(self, rng=None):
            ids = [d for d, q in self.drone_registry.items() if q.quadcopter_type == EntityType.LOYALWINGMAN]
            return ids[int(uniform(0.0, 1.0, 1)[0] * len(ids))] if ids else -1

        old = (np.random.uniform, random.random, refshim.BulletClient.stepSimulation,
               EntitiesManager._select_loyalwingman_randomly, fl_mod.random, lb_mod.random,
               fl_mod.FusedLIDAR.read_data, Level5DumbMultiObs.compute_observation)
        np.random.uniform, random.random = uniform, rnd
        refshim.BulletClient.stepSimulation = stepSimulation
        EntitiesManager._select_loyalwingman_randomly = select_agent
        refshim.Hooks.motor_noise = staticmethod(motor_noise if noise_ratio else (lambda i: np.zeros(4)))
        Level5DumbMultiObjectTask.init_constants = init_constants
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                env = Level5DumbMultiObs(GUI=False)
            em = env.entities_manager
            lws, lms = em.get_all_pursuers(), em.get_all_invaders()
            assert len(lws) == N_LW and len(lms) == N_LM, (len(lws), len(lms))
            slot_of = {q.id: j for j, q in enumerate(lws)}
            slot_of.update({q.id: N_LW + i for i, q in enumerate(lms)})
            drones = lws + lms
            agent_slot = slot_of[em.get_agent().id]
            fr = FuseRandom(seed, env_index, slot_of)
            fl_mod.random = fr; lb_mod.random = fr
            _read, _obs = old[6], old[7]
            fuse_log = {}

            def read_data(self):
                fr.begin(self.parent_id)
                r = _read(self)
                fuse_log[slot_of[self.parent_id]] = list(fr.log)
                return r

            def compute_observation(self):          # zeros(1): the LiDARs are updated by compute_info (:112-150)
                fr.obs_call += 1
                fuse_log.clear()
                return _obs(self)
            fl_mod.FusedLIDAR.read_data = read_data
            Level5DumbMultiObs.compute_observation = compute_observation
            info_box = [{}]
            _info = env.compute_info

            def compute_info():                     # Task.compute_info() at the point Env.step calls compute_info()
                info_box[0] = dict(env.task_progression.compute_info())
                return _info()
            env.compute_info = compute_info

            keys = ("present", "stacked", "mask", "inertial", "teacher_actions", "chosen", "reward", "done", "info", "armed",
                    "pos", "was_reset", "ammo")
            rec = {k: [] for k in keys}

            def snap(info, was_reset):
                armed_lw = [q for q in lws if q.armed]          # get_armed_pursuers(): registry order == slot order
                obs_list, acts = info["student_observations"], info["teacher_actions"]
                assert len(obs_list) == len(acts) == len(armed_lw)
                present = np.zeros(N_LW, dtype=bool)
                stacked = np.ones((N_LW, 6, 3, 13, 26), dtype=np.float32)
                mask = np.zeros((N_LW, 6), dtype=bool)
                inertial = np.zeros((N_LW, 15), dtype=np.float32)
                tact = np.zeros((N_LW, 4), dtype=np.float64)
                chosen = np.full((N_LW, 4, 2), -1, dtype=np.int32)
                for q, o, a in zip(armed_lw, obs_list, acts):
                    j = slot_of[q.id]
                    assert set(o) == {"stacked_spheres", "validity_mask", "inertial_data", "last_action"}
                    present[j] = True
                    stacked[j] = o["stacked_spheres"]; mask[j] = o["validity_mask"]; inertial[j] = o["inertial_data"]
                    assert np.array_equal(o["last_action"], np.asarray(a, dtype=np.float32))
                    tact[j] = a
                    lg = fuse_log.get(j, [])
                    pubs = next((v for k, v in lg if k == "sample"), [])
                    ages = [v for k, v in lg if k == "age"]
                    for i, (p_, a_) in enumerate(zip(pubs, ages)):
                        chosen[j, i] = (p_, a_)
                rec["present"].append(present); rec["stacked"].append(stacked); rec["mask"].append(mask)
                rec["inertial"].append(inertial); rec["teacher_actions"].append(tact); rec["chosen"].append(chosen)
                rec["armed"].append(np.array([q.armed for q in drones]))
                rec["pos"].append(np.array([q.simulation.bodies[q.id].pos for q in drones]))
                rec["ammo"].append([q.gun.munition for q in lws]); rec["was_reset"].append(was_reset)

            obs, info = env.reset()
            assert obs.shape == (1,)
            snap(info, True)
            for t in range(n_steps):
                obs, r, term, trunc, info = env.step(np.zeros(4))
                tinfo = info_box[0]
                rec["reward"].append(r); rec["done"].append(term)
                rec["info"].append([tinfo.get("agent_kills", 0), tinfo.get("allies_kills", 0), tinfo.get("deads", 0),
                                    tinfo.get("current_wave", 0)])
                snap(info, False)
                if term:
                    obs, info = env.reset()
                    snap(info, True)
            out.update({k: np.array(v) for k, v in rec.items()})
            out["counters"] = np.array([ctr["spawn"], ctr["hit"], ctr["phys"], fr.obs_call + 1])
            out["agent_slot"] = np.array(agent_slot)
        finally:
            (np.random.uniform, random.random, refshim.BulletClient.stepSimulation,
             EntitiesManager._select_loyalwingman_randomly, fl_mod.random, lb_mod.random,
             fl_mod.FusedLIDAR.read_data, Level5DumbMultiObs.compute_observation) = old
            Level5DumbMultiObjectTask.init_constants = _init_constants

    th = threading.Thread(target=body)
    th.start(); th.join()
    if not out:
        raise RuntimeError("reference run failed")
    out["meta"] = np.array([seed, env_index, n_steps, 0])
    out["step_increment"] = np.array(100 if step_increment is None else step_increment)
    out["noise_ratio"] = np.array(noise_ratio)
    return out


def run_reference_eval2bt(seed, env_index, n_steps, noise_ratio=0.02, max_step=None):
    """Level52BTEvaluationEnvironment (level5_eval_2bt_environment.py) + Level52BTEvaluationTask: two behaviour-tree
    wingmen vs 5 -> 30 munitions, no observation, no reward; info = kills per drone, deads, current wave
    (apps/threatsense_runner/evaluation_2bt.py).  ``max_step`` overrides the task CONSTANT MAX_STEP (1300, task :112)
    so that a recording of a few hundred steps holds a time-out and a reset.  No logic is patched."""
    N_LW, N_LM = 2, 30
    refshim.install()
    _apply_patches()
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
            slot = N_LW + creation_index if creation_index < N_LM else creation_index - N_LM
            return px.normal4(seed, np.uint32(env_index), np.uint32(ctr["phys"]), np.uint32(slot))

        _step = refshim.BulletClient.stepSimulation

        def stepSimulation(self):
            _step(self)
            ctr["phys"] += 1

        with contextlib.redirect_stdout(io.StringIO()):
            from core.entities.entity_type import EntityType
            from threatsense.level5.components.entities_manager import EntitiesManager
            from threatsense.level5.level5_eval_2bt_environment import Level52BTEvaluationEnvironment
            from threatsense.level5.components.tasks_management.tasks.level5_2bt_evaluation_task import Level52BTEvaluationTask
        _init_constants = Level52BTEvaluationTask.init_constants

        def init_constants(self):
            _init_constants(self)
            if max_step is not None:
                self.MAX_STEP = max_step

        def select_agent(self, rng=None):
            ids = [d for d, q in self.drone_registry.items() if q.quadcopter_type == EntityType.LOYALWINGMAN]
            return ids[int(uniform(0.0, 1.0, 1)[0] * len(ids))] if ids else -1

        old = (np.random.uniform, random.random, refshim.BulletClient.stepSimulation, EntitiesManager._select_loyalwingman_randomly)
        np.random.uniform, random.random = uniform, rnd
        refshim.BulletClient.stepSimulation = stepSimulation
        EntitiesManager._select_loyalwingman_randomly = select_agent
        refshim.Hooks.motor_noise = staticmethod(motor_noise if noise_ratio else (lambda i: np.zeros(4)))
        Level52BTEvaluationTask.init_constants = init_constants
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                env = Level52BTEvaluationEnvironment(GUI=False)
            em = env.entities_manager
            lws, lms = em.get_all_pursuers(), em.get_all_invaders()
            assert len(lws) == N_LW and len(lms) == N_LM, (len(lws), len(lms))
            drones = lws + lms
            assert em.agent_id == -1          # Teacher_Student=False: no agent is ever chosen (no SPAWN draw for it)
            agent_slot = -1
            keys = ("reward", "done", "kills", "info", "armed", "pos", "ammo", "was_reset")
            rec = {k: [] for k in keys}

            def snap(was_reset):
                rec["armed"].append(np.array([q.armed for q in drones]))
                rec["pos"].append(np.array([q.simulation.bodies[q.id].pos for q in drones]))
                rec["ammo"].append([q.gun.munition for q in lws]); rec["was_reset"].append(was_reset)

            obs, info = env.reset()
            assert obs == {} and set(info) == {"kills_per_drone", "deads", "current_wave"}
            snap(True)
            for t in range(n_steps):
                obs, r, term, trunc, info = env.step(np.zeros(4))
                assert obs == {} and r == 0.0 and trunc is False
                rec["reward"].append(r); rec["done"].append(term)
                rec["kills"].append([info["kills_per_drone"][q.id]["kills"] for q in lws])
                rec["info"].append([info["deads"], info["current_wave"]])
                snap(False)
                if term:
                    env.reset()
                    snap(True)
            out.update({k: np.array(v) for k, v in rec.items()})
            out["counters"] = np.array([ctr["spawn"], ctr["hit"], ctr["phys"]])
            out["agent_slot"] = np.array(agent_slot)
        finally:
            (np.random.uniform, random.random, refshim.BulletClient.stepSimulation, EntitiesManager._select_loyalwingman_randomly) = old
            Level52BTEvaluationTask.init_constants = _init_constants

    th = threading.Thread(target=body)
    th.start(); th.join()
    if not out:
        raise RuntimeError("reference run failed")
    out["meta"] = np.array([seed, env_index, n_steps, 0])
    out["noise_ratio"] = np.array(noise_ratio)
    out["max_step"] = np.array(1300 if max_step is None else max_step)
    return out


CASES = [  # (file stem, seed, env_index, steps, policy_seed, noise_ratio, chase_prob, kamikaze_after)
    ("level5_c1_kite", 501, 0, 700, 1, 0.02, 0.9, None),
    ("level5_c1_kite_b", 502, 1, 600, 2, 0.02, 0.9, None),
    ("level5_c1_ram", 503, 4, 400, 3, 0.0, 0.9, 120),
    ("level5_c1_random", 504, 9, 900, 4, 0.02, 0.0, None),
    # Level5FusionEnvironment (base Level5Environment + Level5FusionTask, 6 wingmen vs 5 -> 30 munitions): l5fusion_random
    # holds 245 observations with dead allies on both sides of a living agent (slots 2 and 4, agent 3)
    ("l5fusion_random", 604, 3, 700, 4, 0.02, 0.0, None, "fusion"),
    ("l5fusion_kite", 605, 1, 500, 2, 0.02, 0.9, None, "fusion"),
    ("l5fusion_ram", 603, 0, 300, 1, 0.02, 0.9, 60, "fusion"),
]


DUMB_CASES = [  # (file stem, seed, env_index, steps, noise_ratio, STEP_INCREMENT override)
    ("l5dumb_long", 701, 2, 700, 0.02, None),       # agent slot 0, a wingman dies (step 450+), waves up to 6
    ("l5dumb_timeout", 706, 3, 600, 0.02, 5),       # agent slot 4; 5 extra steps per kill: time-outs and resets
]


EVAL2BT_CASES = [  # (file stem, seed, env_index, steps, noise_ratio, MAX_STEP override)
    ("l5eval2bt_timeout", 801, 1, 400, 0.02, 150),  # two time-outs + resets
    ("l5eval2bt_long", 802, 4, 700, 0.02, None),    # the task's own MAX_STEP = 1300: kills of both wingmen, waves
]


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for stem, seed, env_index, steps, noise, max_step in EVAL2BT_CASES:
        rec = run_reference_eval2bt(seed, env_index, steps, noise, max_step)
        np.savez_compressed(os.path.join(GOLDEN_DIR, stem + ".npz"), **rec)
        print(stem, "episodes:", int(rec["done"].sum()), "kills:", rec["kills"].max(0), "deads, wave:", rec["info"].max(0))
    for stem, seed, env_index, steps, noise, inc in DUMB_CASES:
        rec = run_reference_dumb(seed, env_index, steps, noise, inc)
        np.savez_compressed(os.path.join(GOLDEN_DIR, stem + ".npz"), **rec)
        print(stem, "agent slot:", int(rec["agent_slot"]), "episodes:", int(rec["done"].sum()), "info max:", rec["info"].max(0),
              "observers/step:", float(rec["present"].sum(1).mean()))
    for stem, seed, env_index, steps, pseed, noise, chase, kami, *kind in CASES:
        rec = run_reference(seed, env_index, steps, pseed, noise, chase, kami, *kind)
        np.savez_compressed(os.path.join(GOLDEN_DIR, stem + ".npz"), **rec)
        print(stem, "agent slot:", int(rec["agent_slot"]), "episodes:", int(rec["done"].sum()), "kills:", rec["info"][:, 0].max(),
              "ally kills:", rec["info"][:, 1].max(), "max wave:", rec["info"][:, 3].max(), "counters:", rec["counters"],
              "valid spheres/step:", float(rec["mask"].sum(1).mean()))


if __name__ == "__main__":
    main()

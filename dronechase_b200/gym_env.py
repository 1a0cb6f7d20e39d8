"""Single-env ``gymnasium.Env`` facades with the reference's class names and constructor kwargs.

``Exp02vFinalEnvironment(dome_radius=20, rl_frequency=15, GUI=False)`` & co. mirror
src/threatengage/environments/level4/exp0{2,3,4}_vFinal_environment.py:44-49 so the training
scripts under apps/threatengage_runner keep constructing envs the way they do today; each is a
1-env view of the batched simulator (use ``DroneChaseVecEnv`` for throughput).
"""
from __future__ import annotations

from collections import defaultdict

import numpy as np
import torch

from .config import preset
from .sim import BatchedThreatEngageEnv
from .vec_env import _gym, make_spaces

_EnvBase = _gym.Env if _gym is not None else object


def default_keymap():
    """Env.get_keymap (exp02_vFinal_environment.py:333-349) with plain strings for the keys."""
    key_map = defaultdict(lambda: [0, 0, 0, 1])
    key_map.update({"up": [0, 1, 0, 1], "down": [0, -1, 0, 1], "left": [-1, 0, 0, 1], "right": [1, 0, 0, 1],
                    "e": [0, 0, 1, 1], "d": [0, 0, -1, 1]})
    return key_map


class _Stage03Env(_EnvBase):
    PRESET = "exp02_vFinal"
    SIM_KWARGS: dict = {}

    def __init__(self, dome_radius: float = 20, rl_frequency: int = 15, GUI: bool = False, seed: int = 0, device=0):
        if GUI:
            raise ValueError("the batched GPU simulator has no GUI")
        self.dome_radius, self.rl_frequency = dome_radius, rl_frequency
        self.cfg = preset(self.PRESET, dome_radius=float(dome_radius), rl_frequency=int(rl_frequency))
        self.sim = BatchedThreatEngageEnv(self.cfg, n_envs=1, seed=seed, device=device, auto_reset=False, **self.SIM_KWARGS)
        self.action_space, self.observation_space = make_spaces(self.cfg)
        self.max_step_calls = 20 * rl_frequency

    def _obs(self):
        return {k: v[0].cpu().numpy().copy() for k, v in self.sim.obs.items()}

    def reset(self, seed=0, options=None):
        self.sim.reset()
        return self._obs(), {}

    def step(self, rl_action=np.array([0, 0, 0, 0])):
        a = torch.as_tensor(np.asarray(rl_action, dtype=np.float32)).reshape(1, 4)
        self.sim.step(a.to(self.sim.device))
        info = {k: int(v[0]) for k, v in self.sim.info_dict().items()}
        info = {k: info[k] for k in ("agent_kills", "allies_kills", "deads", "current_wave")}
        return self._obs(), float(self.sim.reward[0]), bool(self.sim.done[0]), False, info

    def close(self):
        self.sim.close()

    def get_keymap(self):
        return default_keymap()


class Exp02vFinalEnvironment(_Stage03Env):
    PRESET = "exp02_vFinal"


class Exp03vFinalEnvironment(_Stage03Env):
    PRESET = "exp03_vFinal"


class Exp04vFinalEnvironment(_Stage03Env):
    PRESET = "exp04_vFinal"


class Exp02V2FullEnvironment(_Stage03Env):
    PRESET = "exp02_v2_full"


class PyflytL3EnviromentV2(_Stage03Env):
    """stage02: threatengage/environments/level3/pyflyt_level3_environment_v2.py:30-36 (dome_radius defaults to 8)."""
    PRESET = "stage02"

    def __init__(self, dome_radius: float = 8, rl_frequency: int = 15, GUI: bool = False, debug_on: bool = False,
                 seed: int = 0, device=0):
        super().__init__(dome_radius=dome_radius, rl_frequency=rl_frequency, GUI=GUI, seed=seed, device=device)

    def step(self, rl_action):
        obs, reward, terminated, truncated, _ = super().step(rl_action)
        return obs, reward, terminated, truncated, {}          # compute_info returns {} (:158-159)


class PyflytL2EnviromentModifiedV2(_Stage03Env):
    """stage01: threatengage/environments/level2/pyflyt_level2_environment_modified_v2.py:27-32 (dome_radius 10)."""
    PRESET = "stage01"

    def __init__(self, dome_radius: float = 10, rl_frequency: int = 15, GUI: bool = False, seed: int = 0, device=0):
        super().__init__(dome_radius=dome_radius, rl_frequency=rl_frequency, GUI=GUI, seed=seed, device=device)

    def step(self, rl_action):
        obs, reward, terminated, truncated, _ = super().step(rl_action)
        return obs, reward, terminated, truncated, {}          # compute_info returns {} (:211-212)


class Level5C1FusionEnvironment(_Stage03Env):
    """threatsense: threatsense/level5/level5_c1_fusion_environment.py:7-9 (``GUI=True, rl_frequency=15``; there is no GUI
    here, the flag is accepted and ignored).  Observation: stacked_spheres (6,3,13,26), validity_mask (6,),
    inertial_data (15,), last_action (4,); ``compute_info`` returns {} (:106-107)."""
    PRESET = "level5_c1"

    def __init__(self, GUI: bool = True, rl_frequency: int = 15, seed: int = 0, device=0):
        super().__init__(dome_radius=20, rl_frequency=rl_frequency, GUI=False, seed=seed, device=device)

    def step(self, rl_action=np.array([0, 0, 0, 0])):
        obs, reward, terminated, truncated, _ = super().step(rl_action)
        return obs, reward, terminated, truncated, {}


class Level5FusionEnvironment(_Stage03Env):
    """threatsense: threatsense/level5/level5_fusion_environment.py:5-16 = the base ``Level5Environment``
    (level5_envrionment.py) with ``Level5FusionTask`` (6 wingmen, 5 -> 30 munitions).  Observation as the base class
    returns it (:296-334): stacked_spheres (6,3,13,26), validity_mask (6,), inertial_data (15,), the env's last_action (4,)
    and the dummy teacher ``lidar`` of zeros (2,13,26).  ``compute_info`` (:291-292) calls compute_observation twice more:
    ``info["teacher_observation"]`` (:336-340: lidar, inertial_data, last_action) and ``info["student_observation"]``
    (:342-346: a second, differently drawn stack of the same ring with its validity mask, inertial_data, last_action --
    ``dc_buffers.student_*``, one more ``stack_kernel`` launch per step)."""
    PRESET = "level5_fusion"
    SIM_KWARGS = {"with_student": True}

    def __init__(self, GUI: bool = True, rl_frequency: int = 15, seed: int = 0, device=0):
        super().__init__(dome_radius=20, rl_frequency=rl_frequency, GUI=False, seed=seed, device=device)

    def _obs(self):
        obs = super()._obs()
        obs["lidar"] = np.zeros((2, 13, 26), dtype=np.float32)
        return obs

    def _info(self, obs):
        student = {k: v[0].cpu().numpy().copy() for k, v in self.sim.student_obs.items()}
        return {"student_observation": student,
                "teacher_observation": {k: obs[k] for k in ("lidar", "inertial_data", "last_action")}}

    def reset(self, seed=0, options=None):
        obs, _ = super().reset(seed=seed, options=options)
        return obs, self._info(obs)

    def step(self, rl_action=np.array([0, 0, 0, 0])):
        obs, reward, terminated, truncated, _ = super().step(rl_action)
        return obs, reward, terminated, truncated, self._info(obs)


class Level5DumbMultiObs(_Stage03Env):
    """threatsense: threatsense/level5/level5_dumb_multiobs.py:9-150 with ``Level5DumbMultiObjectTask`` -- the
    data-collection env of apps/threatsense_runner/collect_and_save.py.  Seven wingmen, all flown by the behaviour tree
    (the action passed to ``step`` is ignored, :98), 5 -> 30 munitions.  The observation is ``zeros(1)`` (:34-35); ``info``
    carries ``student_observations`` (stacked_spheres, validity_mask, inertial_data, last_action of every ARMED wingman)
    and ``teacher_actions`` (their last commands), :112-150."""
    PRESET = "level5_dumb_multiobs"

    def __init__(self, GUI: bool = True, rl_frequency: int = 15, seed: int = 0, device=0):
        super().__init__(dome_radius=20, rl_frequency=rl_frequency, GUI=False, seed=seed, device=device)
        if _gym is not None:
            self.observation_space = _gym.spaces.Box(low=0, high=1, shape=(1,), dtype=np.float32)       # :29-31
            self.action_space = _gym.spaces.Box(low=-1, high=1, shape=(4,), dtype=np.float32)           # :83-85

    def _info(self):
        mo = {k: v[0].cpu().numpy() for k, v in self.sim.multi_obs.items()}
        slots = np.nonzero(mo["present"])[0]
        obs = [{"stacked_spheres": mo["stacked_spheres"][j].copy(), "validity_mask": mo["validity_mask"][j].copy(),
                "inertial_data": mo["inertial_data"][j].copy(), "last_action": mo["last_action"][j].copy()} for j in slots]
        return {"student_observations": obs, "teacher_actions": [mo["last_action"][j].copy() for j in slots]}

    def reset(self, seed=0, options=None):
        self.sim.reset()
        return np.zeros(1, dtype=np.float32), self._info()

    def step(self, action=None):
        self.sim.step(None)
        return np.zeros(1, dtype=np.float32), float(self.sim.reward[0]), bool(self.sim.done[0]), False, self._info()


class Level52BTEvaluationEnvironment(_Stage03Env):
    """threatsense: threatsense/level5/level5_eval_2bt_environment.py:11-77 with ``Level52BTEvaluationTask`` -- the env of
    apps/threatsense_runner/evaluation_2bt.py.  Two behaviour-tree wingmen vs 5 -> 30 munitions; the observation is ``{}``,
    the reward 0.0, ``info`` = kills_per_drone / deads / current_wave (task :470-477; the drones are keyed by wingman slot
    here, the reference keys them by PyBullet body id)."""
    PRESET = "level5_eval_2bt"

    def __init__(self, GUI: bool = True, rl_frequency: int = 15, seed: int = 0, device=0):
        super().__init__(dome_radius=20, rl_frequency=rl_frequency, GUI=False, seed=seed, device=device)
        if _gym is not None:
            self.observation_space = _gym.spaces.Box(low=0, high=1, shape=(1,), dtype=np.float32)       # :24-26
            self.action_space = _gym.spaces.Box(low=-1, high=1, shape=(4,), dtype=np.float32)           # :34-36

    def _info(self):
        i = {k: int(v[0]) for k, v in self.sim.info_dict().items()}
        kills = {0: {"name": "loyalwingman_0", "type": "BT", "kills": i["agent_kills"]},
                 1: {"name": "loyalwingman_1", "type": "BT", "kills": i["allies_kills"]}}
        return {"kills_per_drone": kills, "deads": i["deads"], "current_wave": i["current_wave"]}

    def reset(self, seed=0, options=None):
        self.sim.reset()
        return {}, self._info()

    def step(self, action=None):
        self.sim.step(None)
        return {}, 0.0, bool(self.sim.done[0]), False, self._info()


class Exp05vFinalEnvironment(_Stage03Env):
    """threatengage/environments/level4/exp05_vFinal_environment.py + tasks/exp05_vFinal_task.py: exp03 with the second
    wingman flown by a second policy.  ``update_model(model)`` (:103-105, task :262-263) installs it: an SB3 model (anything
    with ``predict(observation, deterministic=True)``), a ``dronechase_b200.policy.LidarInertialActionPolicy`` or any callable
    on the observation dict of device tensors.  Until a model is installed the reference would raise; here ``step`` does."""
    PRESET = "exp05_vFinal"

    def __init__(self, dome_radius: float = 20, rl_frequency: int = 15, GUI: bool = False, seed: int = 0, device=0):
        super().__init__(dome_radius=dome_radius, rl_frequency=rl_frequency, GUI=GUI, seed=seed, device=device)
        self._drivers = None

    def update_model(self, model):
        from .drivers import TaskDrivers, sb3_policy
        policy = sb3_policy(model) if hasattr(model, "predict") else model
        self._drivers = TaskDrivers(self.sim, {1: policy})

    def reset(self, seed=0, options=None):
        if self._drivers is not None:
            self._drivers.reset()
        return super().reset(seed=seed, options=options)

    def step(self, rl_action=np.array([0, 0, 0, 0])):
        if self._drivers is None:
            raise RuntimeError("Exp05vFinalEnvironment: call update_model(model) first (the second wingman's policy)")
        self._drivers.serve()
        return super().step(rl_action)


class EvaluationEnvironment(_Stage03Env):
    """threatengage/environments/level4/evaluation_environment.py:49-130 + tasks/evaluation_task.py: every wingman is
    flown by the task -- ``configuration["drivers"] = [{"type": "nn", "name": ..., "path": <SB3 zip>}, {"type": "bt", "name":
    ...}, ...]`` (anything else: parked), ``munition_per_defender``, ``ENEMY_BORN_RADIUS``, ``INITIAL_ROUND``,
    ``STEP_INCREMENT``, ``MAX_STEP``, ``TIME_IS_LIMITED`` (:89-110).  ``step`` ignores its argument (:166-183), the reward is
    0, ``info`` = {wingman name: {lw_kills, lw_alive, lw_munitions, current_wave, step}} for the ARMED wingmen (:554-574).
    An "nn" driver is loaded from ``path`` with ``LidarInertialActionPolicy.from_sb3_zip`` (device-resident inference) or
    given directly as ``"policy"`` (a callable on the observation dict of device tensors / an object with ``predict``)."""

    def __init__(self, configuration: dict, dome_radius: float = 20, rl_frequency: int = 15, GUI: bool = False, seed: int = 0, device=0):
        if GUI:
            raise ValueError("the batched GPU simulator has no GUI")
        from .config import evaluation_preset
        from .drivers import TaskDrivers, sb3_policy
        from .policy import LidarInertialActionPolicy
        self.dome_radius, self.rl_frequency = dome_radius, rl_frequency
        self.cfg = evaluation_preset(configuration, dome_radius=float(dome_radius), rl_frequency=int(rl_frequency))
        self.sim = BatchedThreatEngageEnv(self.cfg, n_envs=1, seed=seed, device=device, auto_reset=False)
        self.action_space, self.observation_space = make_spaces(self.cfg)
        self.max_step_calls = 20 * rl_frequency
        self.names = [str(d.get("name", f"lw_{j}")) for j, d in enumerate(configuration["drivers"])]
        policies = {}
        for j in self.cfg.policy_slots:
            d = configuration["drivers"][j]
            p = d.get("policy")
            if p is None:
                p = LidarInertialActionPolicy.from_sb3_zip(d["path"], env=self.sim)
            policies[j] = sb3_policy(p) if hasattr(p, "predict") else p
        self._drivers = TaskDrivers(self.sim, policies) if policies else None

    def _info(self):
        rows = self.sim.lw_info[0].cpu().numpy()
        wave, step = int(self.sim.info[0, 3]), int(self.sim.info[0, 5])
        return {self.names[j]: {"lw_kills": int(rows[j, 0]), "lw_alive": True, "lw_munitions": int(rows[j, 2]),
                                "current_wave": wave, "step": step} for j in range(self.cfg.n_lw) if rows[j, 1]}

    def reset(self, seed=0, options=None):
        self.sim.reset()
        if self._drivers is not None:
            self._drivers.reset()
        return self._obs(), {}

    def step(self, actions_not_used=None):
        if self._drivers is not None:
            self._drivers.serve()
        self.sim.step(None)
        return self._obs(), 0.0, bool(self.sim.done[0]), False, self._info()

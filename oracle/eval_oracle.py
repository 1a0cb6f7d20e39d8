"""CPU oracle -- level4 tasks whose wingmen are flown by policies INSIDE the task (TEST INFRASTRUCTURE, not the product).

  * ``Evaluation_Task`` + ``EvaluationEnvironment`` (src/threatengage/environments/level4/evaluation_environment.py:49-130,
    components/tasks_management/tasks/evaluation_task.py): ``configuration["drivers"]`` names one driver per wingman --
    "nn" (an SB3 PPO, ``driver.predict(compute_lw_observation(pursuer))``), "bt" (LoyalWingmanBehaviorTree) or anything else
    (``pursuer.drive([0,0,0,1])``: zero velocity) (:260-277,630-643); no reward (:512-516), no origin processing (:400),
    the time limit only with TIME_IS_LIMITED (:522), no agent-dead / altitude termination (:518-552), info = per ARMED
    wingman lw_kills / lw_alive / lw_munitions / current_wave / step (:554-574).
  * ``Exp05_vFinal_Task`` + ``Exp05vFinalEnvironment``: exp03 with the second wingman flown by a second policy
    (exp05_vFinal_task.py:252-260, ``update_model``) instead of the behaviour tree.

What is new against oracle/env_oracle.py is the observation a task-driven wingman gets at ``on_step_start``
(``compute_lw_observation``, evaluation_task.py:283-312): ``pursuer.update_lidar()`` runs on the ring as it stands AFTER the
previous step's ``on_step_end`` / ``reset`` -- publishers disarmed there are gone (``MessageHub.terminate``), re-armed ones
have only a slot-0 snapshot and are invisible (STABLE_DELTA_STEP = 1, lidar_buffer.py:285) -- so a wave set-up empties the
sphere of every munition and a reset leaves the previous sphere in place (fused_lidar.py:160-166).  ``ring1[e, d]`` is "d has
a readable slot-1 snapshot".  Every FusedLIDAR keeps ONE sphere that both update points overwrite (``lw_sphere``); the env's
own ``compute_observation`` updates wingman 0's at the end of the step.  ``last_action`` of the observation is the TASK's: the
action of whichever policy-driven wingman was served last (:271), zeroed by ``init_globals`` at a reset (:141).

Pinned by tests/golden/l4eval_*.npz and l4exp05_*.npz (oracle/make_golden_eval.py runs the reference's own classes).
"""
from __future__ import annotations

import dataclasses

import numpy as np

from .env_oracle import EnvOracle, N_PHI, N_THETA, Stage03Config, calculate_rounds, lidar_project, LW_TYPE, LM_TYPE
from .eval_policy import pilot


@dataclasses.dataclass
class DrivenConfig(Stage03Config):
    drivers: tuple = ("agent", "nn_ally")     # per wingman: agent (env action) | nn | nn_ally (exp05: armed pursuers[1:]) | bt | stop
    task: str = "vfinal"                      # "vfinal" (exp05 = exp03's task) | "evaluation"
    time_limited: bool = False                # evaluation: TIME_IS_LIMITED


EXP05 = DrivenConfig(n_lw=2, n_lm=9)


def evaluation_config(drivers, munition=20, born_radius=6.0, initial_round=1, step_increment=100, max_step=300,
                      time_limited=False, **kw) -> DrivenConfig:
    """Evaluation_Task._process_configuration (evaluation_task.py:89-110)."""
    drivers = tuple(drivers)
    return DrivenConfig(n_lw=len(drivers), n_lm=calculate_rounds(len(drivers), munition), munition=munition,
                        born_radius=born_radius, initial_round=initial_round, step_increment=step_increment, max_step=max_step,
                        drivers=drivers, task="evaluation", time_limited=time_limited, **kw)


class DrivenOracle(EnvOracle):
    def __init__(self, cfg: DrivenConfig, n_envs: int, seed: int = 0, env_offset: int = 0, auto_reset: bool = False,
                 salts=None):
        E = n_envs
        self.ring1 = np.zeros((E, cfg.n_drones), dtype=bool)
        self.lw_sphere = np.ones((E, cfg.n_lw, cfg.lidar_channels, N_THETA, N_PHI), dtype=np.float32)
        self.lw_ids = np.full((E, cfg.n_lw, N_THETA, N_PHI), -1, dtype=np.int32)
        self.lw_kills = np.zeros((E, cfg.n_lw), dtype=np.int64)
        self.task_last_action = np.zeros((E, 4), dtype=np.float32)
        self.salts = list(salts) if salts is not None else [0.0] * cfg.n_lw
        self.nn_obs = None                    # filled by step(): what every policy-driven wingman saw / did
        super().__init__(cfg, n_envs, seed=seed, env_offset=env_offset, auto_reset=auto_reset)

    # ------------------------------------------------------------------ ring model
    def _disarm(self, e, d):
        super()._disarm(e, d)
        self.ring1[e, d] = False

    def _reset_env(self, e):
        super()._reset_env(e)
        self.lw_kills[e] = 0
        self.task_last_action[e] = 0

    def _lidar(self, e, obs_slot):
        if not self.ring1[e, obs_slot]:
            return self.lw_sphere[e, obs_slot].copy(), np.full((N_THETA, N_PHI), -1, dtype=np.int32)
        c = self.cfg
        others = [d for d in range(self.D) if d != obs_slot and self.ring1[e, d]]
        types = [LW_TYPE if d < c.n_lw else LM_TYPE for d in others]
        sph, ids = lidar_project(self.imu["position"][e, obs_slot], self.imu["quaternion"][e, obs_slot],
                                 self.imu["position"][e, others], types, others, c.lidar, 2 * c.dome_radius)
        self.lw_sphere[e, obs_slot] = sph
        self.lw_ids[e, obs_slot] = ids
        return sph, ids

    def _inertial(self, e, j):
        c, im = self.cfg, self.imu
        max_speed = 1 * 10 * (1000 / 3600)
        return np.concatenate([np.clip(im["position"][e, j] / c.dome_radius, -1, 1), np.clip(im["velocity"][e, j] / max_speed, -1, 1),
                               np.clip(im["attitude"][e, j] / np.pi, -1, 1), np.clip(im["angular_rate"][e, j] / (2 * np.pi), -1, 1),
                               self._gun_state(e, j)]).astype(np.float32)

    # ------------------------------------------------------------------ pilots
    def _navigate_allies(self, e, allies):
        """drive_lw (evaluation_task.py:257-277) / drive_lw_rl_agent (exp05_vFinal_task.py:252-260)."""
        c = self.cfg
        armed = [j for j in range(c.n_lw) if self.armed[e, j]]
        for j in armed:
            drv = c.drivers[j]
            if drv == "agent":
                continue
            if drv == "nn_ally" and j == armed[0]:
                continue                                   # get_armed_pursuers()[1:]
            if drv in ("nn", "nn_ally"):
                sph, _ = self._lidar(e, j)
                obs = {"lidar": sph.astype(np.float32), "inertial_data": self._inertial(e, j),
                       "last_action": self.task_last_action[e].astype(np.float32)}
                a = pilot(obs["lidar"][None], obs["inertial_data"][None], obs["last_action"][None], self.salts[j])[0]
                self.task_last_action[e] = a
                self.nn_obs["lidar"][e, j] = obs["lidar"]; self.nn_obs["inertial"][e, j] = obs["inertial_data"]
                self.nn_obs["last_action"][e, j] = obs["last_action"]; self.nn_obs["action"][e, j] = a
                self.nn_obs["called"][e, j] = True
                self._drive32(e, j, a)
            elif drv == "bt":
                super()._navigate_allies(e, [j])
            else:
                self._drive(e, j, np.array([0.0, 0.0, 0.0, 1.0]))

    def _drive32(self, e, j, command):
        """Quadcopter.convert_command_to_setpoint (quadcopter.py:379-396) on the float32 array an SB3 policy returns: numpy
        keeps the norm, the division and the product in float32; only the final np.array([...]) is float64."""
        command = np.asarray(command, dtype=np.float32)
        raw = command[:3]
        n = np.linalg.norm(raw)
        direction = raw / (n if n > 0 else 1)
        vx, vy, vz = command[3] * direction
        self.setpoint[e, j] = np.array([vx, vy, 0, vz])

    # ------------------------------------------------------------------ evaluation task
    def _middle(self, e):
        c = self.cfg
        if c.task != "evaluation":
            return super()._middle(e)
        ev = {"shots": [], "explosions": [], "origin": []}
        self._offsets(e)
        shots = 0
        for j, targets in self._in_range(e, c.shoot_range).items():
            if not (self._gun_available(e, j) and self.ammo[e, j] > 0):
                continue
            self.ammo[e, j] -= 1
            self.last_fired[e, j] = self._gun_step(e)
            hit = self._hit_u(e) < c.fire_probability
            ev["shots"].append((j, targets[0], bool(hit)))
            if hit:
                self._disarm(e, targets[0])
                self.lw_kills[e, j] += 1
                shots += 1
        for j, targets in self._in_range(e, c.explosion_range).items():
            self._disarm(e, j); self._disarm(e, targets[0])
            ev["explosions"].append((j, targets[0]))
        if shots > 0:
            self.max_step[e] += c.step_increment
        self.events.append((int(self.step_count[e]), e, ev))
        return 0.0, self._termination(e)

    def _termination(self, e):
        c = self.cfg
        if c.task != "evaluation":
            return super()._termination(e)
        if self._gun_step(e) > self.max_step[e] and c.time_limited:
            return True
        if not self.armed[e, c.n_lw:].any() and self.round[e] >= c.n_lm:
            return True
        if self._outside_dome(e, range(c.n_lw)) > 0:
            return True
        if self._outside_dome(e, range(c.n_lw, self.D)) > 0:
            return True
        return not self.armed[e, :c.n_lw].any()

    # ------------------------------------------------------------------ step
    def _observe(self, after_reset=None):
        obs = super()._observe(after_reset)
        # base class: wingman 0's sphere lives in lidar_obs; it IS that wingman's FusedLIDAR.sphere
        self.lw_sphere[:, 0] = self.lidar_obs
        if self.cfg.task == "evaluation":
            obs["last_action"] = np.zeros((self.E, 4), dtype=np.float32)      # EvaluationEnvironment.last_action is never set
        return obs

    def step(self, actions=None):
        c, E = self.cfg, self.E
        if actions is None:
            actions = np.zeros((E, 4))
        actions = np.asarray(actions, dtype=np.float64)
        self.nn_obs = {"lidar": np.ones((E, c.n_lw, c.lidar_channels, N_THETA, N_PHI), dtype=np.float32),
                       "inertial": np.zeros((E, c.n_lw, 15), dtype=np.float32), "last_action": np.zeros((E, c.n_lw, 4), dtype=np.float32),
                       "action": np.zeros((E, c.n_lw, 4), dtype=np.float32), "called": np.zeros((E, c.n_lw), dtype=bool)}
        self.last_action = actions.copy()
        for e in range(E):
            if c.drivers[0] == "agent":
                self._drive(e, 0, actions[e])
            self._navigate(e)
        self._substeps()
        self.step_count += 1
        self.ring1 = self.armed.copy()                     # AGENT_STEP_BROADCAST: the rings slide, slot 1 = this step's snapshots
        reward = np.zeros(E); done = np.zeros(E, dtype=bool)
        for e in range(E):
            reward[e], done[e] = self._middle(e)
        info = {"agent_kills": self.agent_kills.copy(), "allies_kills": self.allies_kills.copy(), "deads": self.deads.copy(),
                "current_wave": self.round.copy(), "building_life": self.building_life.copy(),
                # Evaluation_Task.compute_info: rows of the ARMED wingmen only
                "lw_kills": self.lw_kills.copy(), "lw_alive": self.armed[:, :c.n_lw].copy(),
                "lw_munitions": self.ammo[:, :c.n_lw].copy(), "step": self.step_count.copy()}
        obs = self._observe()
        self.terminal_obs = obs
        for e in range(E):
            self._step_end(e)
        if self.auto_reset and done.any():
            obs = {k: v.copy() for k, v in obs.items()}
            for e in np.nonzero(done)[0]:
                self._reset_env(e)
            new = self._observe(after_reset=np.ones(E, bool))
            for k in obs:
                obs[k][done] = new[k][done]
        return obs, reward, done, info

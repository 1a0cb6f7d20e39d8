"""CPU oracle -- stage01 ("level2") env step (TEST INFRASTRUCTURE, not the product).

Restates ``PyflytL2EnviromentModifiedV2`` on the shared pieces of oracle/env_oracle.py.  Paths under
/root/reference/src/threatengage/environments/level2:

  ctor / reset / step .... pyflyt_level2_environment_modified_v2.py:27-68, 70-104, 127-146
  reward / termination ... :157-191 (r = -d + 10|v| [d < d_prev] + 1000 [d < 0.4] - 1000 [d > dome];
                           done: steps > 300 or agent or munition outside dome 10)
  catch & teleport ....... :148-155 replace_invader_if_close + components/quadcopter_manager.py:166-179
                           replace_invader: setpoint (x, y, 0, z) in QuadX mode 7, followed by ONE extra
                           update_imu / update_control / update_physics whose force stays applied until the
                           next stepSimulation (it adds to the first substep of the next env step)
  simulation ............. components/pyflyt_level2_simulation.py:87-105: every drone is always stepped
  drones ................. agent wingman (mode 6, RL), idle wingman (mode 6, zero setpoint), munition (mode 7)

Slots: [agent, idle wingman, munition].  The level2 env never publishes AGENT_STEP_BROADCAST, so at HEAD its
refactored LiDAR ring never slides and the sphere stays empty; the harness (oracle/make_golden_stage01.py,
patch P6) adds the broadcast the other levels have, and this oracle follows that intended behaviour.
"""
from __future__ import annotations

import dataclasses

import numpy as np

from . import dynamics as dy
from . import philox as px
from .env_oracle import EnvOracle, Stage03Config
from .stage02_oracle import NO_GROUND

STAGE01 = Stage03Config(n_lw=2, n_lm=1, munition=0, dome_radius=10.0, max_step=300, initial_round=1)
CATCH_DISTANCE = 0.4
AGENT, IDLE, LM = 0, 1, 2


class Stage01Oracle(EnvOracle):
    def __init__(self, cfg: Stage03Config = STAGE01, n_envs: int = 1, seed: int = 0, env_offset: int = 0,
                 auto_reset: bool = False):
        self.pending_f = np.zeros((n_envs, 3, 3)); self.pending_t = np.zeros((n_envs, 3, 3))
        self.last_distance = np.full(n_envs, 2 * cfg.dome_radius)
        self.mode = np.array([6, 6, 7])
        self.caught = np.zeros(n_envs, dtype=bool)
        super().__init__(cfg, n_envs, seed=seed, env_offset=env_offset, auto_reset=auto_reset)
        self.prm = dy.QuadParams(noise_ratio=cfg.noise_ratio, ground_z=NO_GROUND)

    def _u3(self, e):
        return -1.0 + (1.0 - -1.0) * self._spawn_u(e, 3)          # np.random.uniform(-1, 1, 3)

    def _noise(self, e, d):
        if self.prm.noise_ratio == 0.0:
            return np.zeros(4)
        return px.normal4(self.seed, self.env_ids[e], np.uint32(self.phys_ctr[e]), np.uint32(d))

    def _replace_invader(self, e, position):
        """quadcopter_manager.py:166-179."""
        prm = dy.QuadParams(noise_ratio=self.cfg.noise_ratio, ground_z=NO_GROUND)
        self._replace(e, LM, position)
        self.setpoint[e, LM] = [position[0], position[1], 0.0, position[2]]
        self._update_imu(e, LM)
        imu = {k: v[e, LM] for k, v in self.imu.items()}
        pid = self.pid[e, LM].copy()
        pwm = dy.control_update(pid, imu, self.setpoint[e, LM], 7, prm)
        self.pid[e, LM] = pid
        thr, f, t = dy.actuate(self.throttle[e, LM], pwm, imu["velocity"], self._noise(e, LM), prm)
        self.throttle[e, LM] = thr
        self.pending_f[e, LM] += f; self.pending_t[e, LM] += t

    def _env_init(self, e):
        p = self._u3(e)
        for d, pos in ((LM, p), (AGENT, -p), (IDLE, np.array([3.0, 3.0, 3.0]))):
            self.pos[e, d] = pos; self.formation[e, d] = pos; self.armed[e, d] = True; self._update_imu(e, d)
        self.ammo[e] = 0
        self.setpoint[e, LM] = 0          # the spawn-time drive() never survives the reset that follows

    def _reset_env(self, e):
        self.step_count[e] = 0
        self.last_action[e] = 0
        self.agent_kills[e] = 0            # catches of the episode (our counter; the reference's info is {})
        self.last_distance[e] = 2 * self.cfg.dome_radius
        self._replace_invader(e, self._u3(e))
        self._replace(e, AGENT, self._u3(e)); self._update_imu(e, AGENT)
        self._replace(e, IDLE, self._u3(e)); self._update_imu(e, IDLE)
        self._update_last_distance(e)

    def _distance(self, e):
        return float(np.linalg.norm(self.imu["position"][e, LM] - self.imu["position"][e, AGENT]))

    def _update_last_distance(self, e):
        self.last_distance[e] = self._distance(e)

    def _navigate(self, e):
        pass                               # the idle wingman and the munition keep their setpoints

    def _substeps(self):
        c, prm = self.cfg, self.prm
        for _ in range(c.substeps):
            s = dy.imu_state(self.pos, self.quat, self.vel, self.omega)
            s["quaternion"] = dy.quat_from_euler(s["attitude"])
            self.imu = s
            pwm = dy.control_update(self.pid, s, self.setpoint, np.broadcast_to(self.mode, (self.E, 3)), prm)
            if prm.noise_ratio != 0.0:
                sub = np.broadcast_to(np.arange(self.D, dtype=np.uint32), (self.E, self.D))
                noise = px.normal4(self.seed, self.env_ids[:, None], self.phys_ctr[:, None].astype(np.uint32), sub)
            else:
                noise = np.zeros((self.E, self.D, 4))
            self.throttle, f, t = dy.actuate(self.throttle, pwm, s["velocity"], noise, prm)
            f = f + self.pending_f; t = t + self.pending_t
            self.pending_f[:] = 0; self.pending_t[:] = 0
            self.pos, self.quat, self.vel, self.omega = dy.rigid_body_step(self.pos, self.quat, self.vel, self.omega, f, t, prm)
            self.phys_ctr += 1

    def _gun_state(self, e, j):
        return np.array([0.0, 0.0, 1.0])   # set_munition(0): [0/1, 0, available] (gun.py:49-54,68-71,101-113)

    def _middle(self, e):
        c = self.cfg
        d = self._distance(e)
        bonus = penalty = 0.0
        self.reward_margin[e] = min(self.reward_margin[e], abs(d - self.last_distance[e]))
        if d < self.last_distance[e]:
            bonus += 10 * np.linalg.norm(self.imu["velocity"][e, AGENT])
        self.min_margin[e] = min(self.min_margin[e], abs(d - CATCH_DISTANCE), abs(d - c.dome_radius))
        if d < CATCH_DISTANCE:
            bonus += 1000
        if d > c.dome_radius:
            penalty += 1000
        reward = -d + bonus - penalty
        done = bool(self.step_count[e] > c.max_step)
        for slot in (AGENT, LM):
            n0 = float(np.linalg.norm(self.imu["position"][e, slot]))
            self.min_margin[e] = min(self.min_margin[e], abs(n0 - c.dome_radius))
            done |= n0 > c.dome_radius
        self.caught[e] = d < CATCH_DISTANCE
        return reward, done

    def _step_end(self, e):
        if self.caught[e]:
            self.agent_kills[e] += 1
            self._replace_invader(e, self._u3(e))
        self._update_last_distance(e)

"""CPU oracle -- stage02 ("level3") env step (TEST INFRASTRUCTURE, not the product).

Restates ``PyflytL3EnviromentV2`` + ``L3Stage1`` on top of the shared pieces of
oracle/env_oracle.py (dynamics, gun, projection LiDAR, observation vector).  Paths
under /root/reference/src/threatengage/environments/level3:

  step order ............. pyflyt_level3_environment_v2.py:125-152
  episode / reset ........ components/stages.py:96-135 (on_env_init, on_reset, on_episode_start/end)
  pilots ................. components/quadcopter_manager.py:176-205: armed munitions are driven with
                           [0,0,0,0.5] (zero direction -> they hover); drive_support_pursuers iterates the
                           INVADERS again, so the support wingman keeps its zero setpoint
  engagement ............. stages.py:186-215 + quadcopter_manager.py:155-169: a wingman with an empty gun
                           "shoots" successfully without a roll; explode range 0.2 on the same distances
  respawn ................ stages.py:170-174,370-376: disarmed munitions reappear on r in U(2,6) every step
  reward ................. stages.py:234-296 with OffsetHandler.current/last_closest_pursuer_to_invader_distance
                           components/offsets_handler.py:56-66,96-135,186-200 (last offsets filtered to the
                           invaders present now, row 0 = first armed pursuer)
  termination ............ stages.py:298-344 (step > 600, outside dome 8, fewer than 2 armed wingmen)
  positions .............. stages.py:350-368 generate_positions(n, r, r_max): radius, theta, phi draws

Pinned by tests/golden/stage02_*.npz (oracle/make_golden_stage02.py runs the reference's own classes).
"""
from __future__ import annotations

import dataclasses

import numpy as np

from .env_oracle import EnvOracle, Stage03Config
from . import dynamics as dy

STAGE02 = Stage03Config(n_lw=2, n_lm=5, munition=4, dome_radius=8.0, max_step=600, initial_round=5,
                        born_radius=2.0, lw_spawn_radius=1.0, lm_speed=0.5)
NO_GROUND = -1.0e9          # level3 spawns no plane (pyflyt_level3_simulation.py:63-81)


class Stage02Oracle(EnvOracle):
    SUPPORT_MUNITION = 10   # Gun() default (gun.py:11); only the agent gets set_munition(4)
    RESPAWN_R = (2.0, 6.0)

    def __init__(self, cfg: Stage03Config = STAGE02, n_envs: int = 1, seed: int = 0, env_offset: int = 0,
                 auto_reset: bool = False):
        self.last_dist = None
        self.mid_armed = np.zeros((n_envs, cfg.n_drones), dtype=bool)
        self.use_mid = False
        super().__init__(cfg, n_envs, seed=seed, env_offset=env_offset, auto_reset=auto_reset)
        self.prm = dy.QuadParams(noise_ratio=cfg.noise_ratio, ground_z=NO_GROUND)

    # generate_positions(n, r, r_max) stages.py:350-368 (radius draws, then thetas, then phis)
    def gen3(self, e, n, r, r_max=0.0):
        if r > r_max:
            r_max = r
        radius = r + (r_max - r) * self._spawn_u(e, n)
        thetas = 0.0 + (2 * np.pi - 0.0) * self._spawn_u(e, n)
        phis = 0.0 + (np.pi / 2 - 0.0) * self._spawn_u(e, n)
        return np.column_stack((radius * np.sin(phis) * np.cos(thetas), radius * np.sin(phis) * np.sin(thetas),
                                radius * np.cos(phis)))

    def _arm(self, e, d):
        super()._arm(e, d)
        c = self.cfg
        self.ammo[e, d] = c.munition if d == 0 else self.SUPPORT_MUNITION

    def _max_munition(self, j):
        return self.cfg.munition if j == 0 else self.SUPPORT_MUNITION

    def _env_init(self, e):
        c = self.cfg
        if self.last_dist is None:
            self.last_dist = np.full((self.E, c.n_lm), np.nan)
        lm = self.gen3(e, c.n_lm, 2.0)
        for i in range(c.n_lm):
            d = c.n_lw + i
            self.pos[e, d] = lm[i]; self.formation[e, d] = lm[i]; self.armed[e, d] = True; self._update_imu(e, d)
        lw = self.gen3(e, c.n_lw, 1.0)
        for j in range(c.n_lw):
            self.pos[e, j] = lw[j]; self.formation[e, j] = lw[j]; self.armed[e, j] = True; self._update_imu(e, j)
        self._episode_start(e)

    def _episode_start(self, e):
        c = self.cfg
        for d in range(self.D):
            self._arm(e, d)
        self._offsets(e)
        self._remember_distances(e)

    def _reset_env(self, e):
        c = self.cfg
        self.last_action[e] = 0
        self.step_count[e] = 0
        self.agent_kills[e] = self.allies_kills[e] = self.deads[e] = 0
        for d in range(self.D):
            self._disarm(e, d)
        lm = self.gen3(e, c.n_lm, *self.RESPAWN_R)
        for i in range(c.n_lm):
            self._replace(e, c.n_lw + i, lm[i])
        lw = self.gen3(e, c.n_lw, 1.0)
        for j in range(c.n_lw):
            self._replace(e, j, lw[j])
        self._episode_start(e)

    def _navigate(self, e):
        c = self.cfg
        if self.armed[e, :c.n_lw].any():
            for d in range(c.n_lw, self.D):
                if self.armed[e, d]:
                    self._drive(e, d, np.array([0.0, 0.0, 0.0, 0.5]))

    def _row0(self, e):
        """distances[0]: row of the first armed pursuer of the snapshot."""
        for j in range(self.cfg.n_lw):
            if self.off_armed[e, j]:
                return j
        return -1

    def _remember_distances(self, e):
        c = self.cfg
        j = self._row0(e)
        for i in range(c.n_lm):
            d = c.n_lw + i
            self.last_dist[e, i] = (np.linalg.norm(self.off_pos[e, j] - self.off_pos[e, d])
                                    if j >= 0 and self.off_armed[e, d] else np.nan)

    def _gun_state(self, e, j):
        c = self.cfg
        wait = max(c.cooldown_steps - (self.step_count[e] - self.last_fired[e, j]), 0)
        mx = self._max_munition(j)
        return np.array([self.ammo[e, j] / (mx if mx > 0 else 1), wait / c.cooldown_steps,
                         int(self._gun_available(e, j))])

    def _middle(self, e):
        c = self.cfg
        ev = {"shots": [], "explosions": [], "origin": []}
        self._offsets(e)
        shots = 0
        for j, targets in self._in_range(e, c.shoot_range).items():
            if self.ammo[e, j] == 0:                   # "LW suicided to kill LM" quadcopter_manager.py:158-161
                self._disarm(e, targets[0]); shots += 1
                ev["shots"].append((j, targets[0], True)); continue
            if not (self._gun_available(e, j) and self.ammo[e, j] > 0):
                continue
            self.ammo[e, j] -= 1
            self.last_fired[e, j] = self.step_count[e]
            hit = self._hit_u(e) < c.fire_probability
            ev["shots"].append((j, targets[0], bool(hit)))
            if hit:
                self._disarm(e, targets[0]); shots += 1
        exploded = 0
        for j, targets in self._in_range(e, c.explosion_range).items():
            self._disarm(e, j); self._disarm(e, targets[0]); exploded += 1
            ev["explosions"].append((j, targets[0]))
        self.agent_kills[e] += shots; self.deads[e] += exploded
        # ---- reward stages.py:234-296
        munition, reload_progress, gun_available = self._gun_state(e, 0)
        j0 = self._row0(e)
        cur = [np.linalg.norm(self.off_pos[e, j0] - self.off_pos[e, d]) for d in range(c.n_lw, self.D) if self.off_armed[e, d]]
        current = min(cur)
        last = min(self.last_dist[e, i] for i in range(c.n_lm)
                   if self.off_armed[e, c.n_lw + i] and not np.isnan(self.last_dist[e, i]))
        if gun_available == 1 or munition == 0:
            score = -current
        else:
            score = current * (2 * reload_progress - 1)
        bonus = penalty = 0.0
        self.reward_margin[e] = min(self.reward_margin[e], abs(last - current - 0.01))
        if 0.01 < last - current and (gun_available == 1 or munition == 0):
            bonus += 10 * np.linalg.norm(self.imu["velocity"][e, 0])
        bonus += 1000 * shots
        penalty += 1000 * exploded
        if self._outside_dome(e, range(c.n_lw)) > 0:
            penalty += 1000
        reward = score + bonus - penalty
        # ---- termination stages.py:298-344
        done = bool(self.step_count[e] > c.max_step)
        done |= self._outside_dome(e, range(c.n_lw)) > 0
        done |= self._outside_dome(e, range(c.n_lw, self.D)) > 0
        done |= int(self.armed[e, :c.n_lw].sum()) < c.n_lw
        # ---- disarmed munitions reappear stages.py:170-174
        self.mid_armed[e] = self.armed[e]
        dead = [d for d in range(c.n_lw, self.D) if not self.armed[e, d]]
        if dead:
            pos = self.gen3(e, len(dead), *self.RESPAWN_R)
            for d, p in zip(dead, pos):
                self._replace(e, d, p)
            for d in dead:
                self._arm(e, d)
        self.events.append((int(self.step_count[e]), e, ev))
        return reward, done

    def _lidar(self, e, obs_slot):
        # a munition re-armed inside on_step_middle publishes into ring slot 0 only (lidar_buffer.py:401-437),
        # the sphere reads slot 1: it shows up one step later
        if not self.use_mid:
            return super()._lidar(e, obs_slot)
        keep = self.armed[e].copy()
        self.armed[e] = self.mid_armed[e]
        out = super()._lidar(e, obs_slot)
        self.armed[e] = keep
        return out

    def _step_end(self, e):
        self._remember_distances(e)          # OffsetHandler.on_end_step: last := current (pre-respawn positions)

    def step(self, actions):
        self.use_mid = True
        try:
            return super().step(actions)
        finally:
            self.use_mid = False

"""CPU oracle -- batched restatement of the stage03 ("level4") env step.

TEST INFRASTRUCTURE, not the product: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg may import this.  numpy float64, per-env Python loops
for the order-dependent game logic, vectorised (E, D) arrays for the dynamics.

Pinned against the reference's own code: oracle/make_golden.py executes
``Exp02vFinalEnvironment`` & co. from /root/reference/src (through oracle/refshim)
with the same Philox streams and records trajectories under tests/golden/;
tests/test_oracle_golden.py replays them through this file.

What each block follows (paths under /root/reference/src):
  step order ............ threatengage/environments/level4/exp02_vFinal_environment.py:155-188
  16 substeps ........... .../level4/components/simulation/level4_simulation.py:84-98
  command -> setpoint ... core/entities/quadcopters/quadcopter.py:379-413
  arm/disarm/replace .... core/entities/quadcopters/quadcopter.py:433-482
  LM pursuit FSM ........ core/entities/navigators/loitering_munition_navigator.py:51-230
                          core/entities/navigators/loitering_munition_navigator_air_combat_only.py:83-105
  LW behaviour tree ..... core/entities/navigators/loyalwingman_navigator.py:32-86,238-352
  offsets snapshot ...... core/context/offsets_handler.py:38-95,167-309,341-391
  gun ................... core/entities/quadcopters/components/weapons/gun.py:8-124
  engagement/reward/termination/waves
                          .../level4/components/tasks_management/tasks/exp02_vFinal_task.py:87-664
                          .../tasks/exp02_v2_full_task.py (protected-area variant)
  fused LiDAR ........... core/entities/quadcopters/components/sensors/fused_lidar.py:143-217
                          .../sensors/components/lidar_math.py:25-34,53-96,128-129,274-311
  classic LiDAR ......... core/entities/quadcopters/components/sensors/lidar.py:151-200,263-319
  observation ........... exp02_vFinal_environment.py:206-234; .../utils/normalization.py:6-110

Drone slots per env: [LW_0 (the RL agent), LW_1.., LM_0, LM_1, ...].
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import dynamics as dy
from . import philox as px

NAV_WAIT, NAV_WINGMAN, NAV_BUILDING = 0, 1, 2
LM_TYPE, LW_TYPE = 1, 3                   # EntityType values (entity_type.py)
N_THETA, N_PHI = 13, 26                   # LIDARSpec(resolution=16) (angle_grid.py:29-40)


@dataclass
class Stage03Config:
    """One POD for the whole task family; mirrored by dc_config in include/."""
    n_lw: int = 1
    n_lm: int = 6                         # = MAX_NUMBER_OF_ROUNDS (calculate_rounds)
    munition: int = 20
    dome_radius: float = 20.0
    born_radius: float = 6.0              # ENEMY_BORN_RADIUS
    lw_spawn_radius: float = 2.0
    explosion_range: float = 0.2
    shoot_range: float = 1.0
    step_increment: int = 100
    max_step: int = 300
    initial_round: int = 1
    cooldown_steps: float = 60.0          # 4 s / (1/15 s)  (gun.py:12-13,25)
    fire_probability: float = 0.9
    lm_speed: float = 0.4
    bt_speed: float = 0.6
    lm_nav: str = "air"                   # "air" (air_combat_only) | "full" (cone FSM)
    ally_mode: str = "bt"                 # "bt" | "stop" (constant command)
    ally_stop_mag: float = 1.0
    reward: str = "vfinal"                # "vfinal" | "v2full"
    vel_bonus: float = 1.0                # 10 in exp04
    building: tuple = (0.0, 0.0, 0.1)
    fixed_lw_spawn: bool = False          # exp02_v2_full re-uses the init positions
    lidar: str = "fused"                  # "fused" (3 ch) | "classic" (2 ch)
    noise_ratio: float = 0.02
    substeps: int = 16                    # int(120/15) * (240//120)

    @property
    def n_drones(self):
        return self.n_lw + self.n_lm

    @property
    def lidar_channels(self):
        return 3 if self.lidar == "fused" else 2


def calculate_rounds(num_defenders, munition_per_defender):
    """exp02_vFinal_task.py:197-225."""
    total = num_defenders * munition_per_defender
    return math.ceil((-1 + math.sqrt(1 + 8 * total)) / 2)


PRESETS = {
    "exp02_vFinal": Stage03Config(),
    "exp03_vFinal": Stage03Config(n_lw=2, n_lm=9),
    "exp04_vFinal": Stage03Config(n_lw=2, n_lm=9, ally_mode="stop", ally_stop_mag=1.0, vel_bonus=10.0),
    "exp02_v2_full": Stage03Config(born_radius=8.0, lm_nav="full", reward="v2full",
                                   ally_mode="stop", ally_stop_mag=0.5, fixed_lw_spawn=True),
    "swarm": Stage03Config(n_lw=4, n_lm=64, initial_round=64),
}


def lidar_project(own_pos, own_quat, ent_pos, ent_type, ent_id, flavour="fused", radius=40.0):
    """Projection "LiDAR": known entity centres -> 13x26 spherical grid, nearest wins.

    fused   : FusedLIDAR.update_data (fused_lidar.py:143-217) + LidarMath.reframe /
              cartesian_to_spherical / index_from_radian / add_features
              (lidar_math.py:25-34,53-96,128-129,274-311).  Snapshot position and
              quaternion are float32 (perception_snapshot.py:91-110), the math float64,
              index by truncation, strict '<' so the first entity wins a tie.
    classic : LIDAR.update_data/_add_end_position/_add_spherical/_normalize_angle
              (lidar.py:151-200,263-308).  float64 inputs, cull unless 0 < r < radius,
              index by Python round() modulo n, '>' rejects so the last entity wins a tie.
    Returns (sphere float32 (C,13,26), ids int32 (13,26) with the winning entity id or -1).
    """
    fused = flavour == "fused"
    sphere = np.ones((3 if fused else 2, N_THETA, N_PHI), dtype=np.float32)
    ids = np.full((N_THETA, N_PHI), -1, dtype=np.int32)
    if fused:
        own_p = np.asarray(own_pos).astype(np.float32)
        own_q = np.asarray(own_quat).astype(np.float32)
        nsq = np.dot(own_q, own_q)
        qinv = np.array([-own_q[0], -own_q[1], -own_q[2], own_q[3]], dtype=np.float32) / nsq
        Rinv = dy.rot_from_quat(qinv.astype(np.float64))
    else:
        own_p = np.asarray(own_pos, dtype=np.float64)
        Rinv = dy.rot_from_quat(np.asarray(own_quat, dtype=np.float64)).T
    for p, etype, eid in zip(ent_pos, ent_type, ent_id):
        if fused:
            p = np.asarray(p).astype(np.float32)
            rel = Rinv @ (p.astype(np.float64) - own_p.astype(np.float64))
        else:
            rel = Rinv @ (np.asarray(p, dtype=np.float64) - own_p)
        x, y, z = rel
        r = np.sqrt(x * x + y * y + z * z)
        if r == 0:
            theta = phi = 0.0
        else:
            theta = np.arccos(np.clip(z / r, -1.0, 1.0)); phi = np.arctan2(y, x)
        if fused:
            rn = np.clip(r / radius, 0, 1)
            ti = int(np.clip(int(theta / np.pi * N_THETA), 0, N_THETA - 1))
            pj = int(np.clip(int((phi + np.pi) / (2 * np.pi) * N_PHI), 0, N_PHI - 1))
            if rn < sphere[0, ti, pj]:
                sphere[0, ti, pj] = rn; sphere[1, ti, pj] = etype / 5; sphere[2, ti, pj] = 0.1
                ids[ti, pj] = eid
        else:
            if not (r > 0 and r < radius):
                continue
            rn = r / radius
            ti = round(theta / np.pi * N_THETA) % N_THETA
            pj = round((phi + np.pi) / (2 * np.pi) * N_PHI) % N_PHI
            if rn > sphere[0, ti, pj]:
                continue
            sphere[0, ti, pj] = rn; sphere[1, ti, pj] = etype / 5
            ids[ti, pj] = eid
    return sphere, ids


class EnvOracle:
    def __init__(self, cfg: Stage03Config, n_envs: int, seed: int = 0, env_offset: int = 0,
                 auto_reset: bool = False):
        self.cfg = cfg
        self.E, self.D = n_envs, cfg.n_drones
        self.seed = seed
        self.env_ids = np.arange(env_offset, env_offset + n_envs, dtype=np.uint32)
        self.auto_reset = auto_reset
        self.prm = dy.QuadParams(noise_ratio=cfg.noise_ratio)
        E, D = self.E, self.D
        self.pos = np.zeros((E, D, 3)); self.quat = np.zeros((E, D, 4)); self.quat[..., 3] = 1
        self.vel = np.zeros((E, D, 3)); self.omega = np.zeros((E, D, 3))
        self.throttle = np.zeros((E, D, 4)); self.pid = np.zeros((E, D, dy.PID_WORDS))
        self.setpoint = np.zeros((E, D, 4))
        self.armed = np.zeros((E, D), dtype=bool)
        self.ammo = np.zeros((E, D), dtype=np.int64)
        self.last_fired = np.full((E, D), -cfg.cooldown_steps)
        self.formation = np.zeros((E, D, 3))
        self.imu = {"position": np.zeros((E, D, 3)), "attitude": np.zeros((E, D, 3)),
                    "velocity": np.zeros((E, D, 3)), "angular_rate": np.zeros((E, D, 3)),
                    "quaternion": np.zeros((E, D, 4))}
        self.imu["quaternion"][..., 3] = 1
        # offsets snapshot (OffsetHandler.current_offsets)
        self.off_armed = np.zeros((E, D), dtype=bool)
        self.off_pos = np.zeros((E, D, 3))
        self.nav = np.zeros((E, D), dtype=np.int64)
        self.step_count = np.zeros(E, dtype=np.int64)
        self.max_step = np.full(E, cfg.max_step, dtype=np.int64)
        self.round = np.full(E, cfg.initial_round, dtype=np.int64)
        self.agent_kills = np.zeros(E, dtype=np.int64); self.allies_kills = np.zeros(E, dtype=np.int64)
        self.deads = np.zeros(E, dtype=np.int64); self.building_life = np.ones(E, dtype=np.int64)
        self.last_closest = np.full(E, cfg.dome_radius)
        self.hit_ctr = np.zeros(E, dtype=np.int64); self.spawn_ctr = np.zeros(E, dtype=np.int64)
        self.phys_ctr = np.zeros(E, dtype=np.int64)
        self.last_action = np.zeros((E, 4))
        self.lw_init_pos = np.zeros((E, cfg.n_lw, 3))
        # FusedLIDAR.sphere survives a reset untouched (fused_lidar.py:160-166: update_data
        # returns early while the ring buffer is empty), so the sphere is persistent state.
        self.lidar_obs = np.ones((E, cfg.lidar_channels, N_THETA, N_PHI), dtype=np.float32)
        self.lidar_ids = np.full((E, N_THETA, N_PHI), -1, dtype=np.int32)
        self.events = []                    # per-step engagement event log (tests)
        self.min_margin = np.full(E, np.inf)  # distance of any event predicate from its threshold
        self.reward_margin = np.full(E, np.inf)  # same for the reward-only "got closer" predicate
        for e in range(E):
            self._env_init(e)

    # ------------------------------------------------------------------ random
    def _spawn_u(self, e, n):
        idx = self.spawn_ctr[e] + np.arange(n)
        self.spawn_ctr[e] += n
        return px.uniform(self.seed, self.env_ids[e], px.STREAM_SPAWN, idx.astype(np.uint32))

    def _hit_u(self, e):
        u = px.uniform(self.seed, self.env_ids[e], px.STREAM_HIT, np.uint32(self.hit_ctr[e]))
        self.hit_ctr[e] += 1
        return float(u)

    def generate_positions(self, e, n, r, min_z=4.0):
        """exp02_vFinal_task.py:583-607 (thetas first, then phis)."""
        thetas = 0.0 + (np.pi - 0.0) * self._spawn_u(e, n)
        lower = min(min_z, r)
        min_phi = np.arccos(lower / r)
        lo = min_phi if r >= min_z else 0.0
        phis = lo + (np.pi / 2 - lo) * self._spawn_u(e, n)
        xs = r * np.sin(phis) * np.cos(thetas)
        ys = r * np.sin(phis) * np.sin(thetas)
        zs = r * np.cos(phis)
        return np.column_stack((xs, ys, zs))

    # ----------------------------------------------------------- entity events
    def _update_imu(self, e, d):
        s = dy.imu_state(self.pos[e, d], self.quat[e, d], self.vel[e, d], self.omega[e, d])
        s["quaternion"] = dy.quat_from_euler(s["attitude"])        # imu.py:38
        for k, v in s.items():
            self.imu[k][e, d] = v

    def _disarm(self, e, d):
        """Quadcopter.disarm (quadcopter.py:461-478)."""
        self.vel[e, d] = 0; self.omega[e, d] = 0
        self.armed[e, d] = False
        self.throttle[e, d] = 0; self.setpoint[e, d] = 0

    def _arm(self, e, d):
        """Quadcopter.arm (quadcopter.py:445-459): imu refresh + gun.reset()."""
        self.armed[e, d] = True
        self._update_imu(e, d)
        self.ammo[e, d] = self.cfg.munition if d < self.cfg.n_lw else 10
        self.last_fired[e, d] = -self.cfg.cooldown_steps

    def _replace(self, e, d, position):
        """Quadcopter.replace (quadcopter.py:433-439): teleport, zero velocity."""
        self.pos[e, d] = position
        self.quat[e, d] = (0, 0, 0, 1)
        self.vel[e, d] = 0; self.omega[e, d] = 0
        self.formation[e, d] = position
        if self.armed[e, d]:
            self._update_imu(e, d)

    def _offsets(self, e):
        """OffsetHandler.calculate_invader_offsets_from_pursuers (offsets_handler.py:68-95)."""
        self.off_armed[e] = self.armed[e]
        self.off_pos[e] = self.imu["position"][e]

    # ----------------------------------------------------------------- episodes
    def _env_init(self, e):
        """Task.on_env_init + on_episode_start in Env.__init__ (exp02_vFinal_environment.py:62-63)."""
        c = self.cfg
        lm = self.generate_positions(e, c.n_lm, c.born_radius)
        for i in range(c.n_lm):
            d = c.n_lw + i
            self.pos[e, d] = lm[i]; self.formation[e, d] = lm[i]
            self.armed[e, d] = True; self._update_imu(e, d)
        for i in range(1, c.n_lm):
            self._disarm(e, c.n_lw + i)
        lw = self.generate_positions(e, c.n_lw, c.lw_spawn_radius)
        self.lw_init_pos[e] = lw
        for j in range(c.n_lw):
            self.pos[e, j] = lw[j]; self.formation[e, j] = lw[j]
            self.armed[e, j] = True; self._update_imu(e, j)
            self.ammo[e, j] = c.munition
        self._episode_start(e)

    def _setup_round(self, e, k):
        """exp02_vFinal_task.py:179-195."""
        c = self.cfg
        for i in range(c.n_lm):
            self._disarm(e, c.n_lw + i)
        positions = self.generate_positions(e, k, c.born_radius)
        for i in range(k):
            self._replace(e, c.n_lw + i, positions[i])
            self._arm(e, c.n_lw + i)

    def _episode_start(self, e):
        """exp02_vFinal_task.py:258-267."""
        c = self.cfg
        self.round[e] = c.initial_round
        self._setup_round(e, c.initial_round)
        for j in range(c.n_lw):
            self._arm(e, j)
        lw = self.lw_init_pos[e] if c.fixed_lw_spawn else self.generate_positions(e, c.n_lw, c.lw_spawn_radius)
        for j in range(c.n_lw):
            self._replace(e, j, lw[j])
        self._offsets(e)
        self.nav[e] = NAV_WAIT

    def _reset_env(self, e):
        """Env.reset (exp02_vFinal_environment.py:133-151) -> task.on_reset (:254-273)."""
        c = self.cfg
        self.last_action[e] = 0
        self.step_count[e] = 0
        self.max_step[e] = c.max_step
        self.agent_kills[e] = self.allies_kills[e] = self.deads[e] = 0
        self.building_life[e] = 1
        self.last_closest[e] = c.dome_radius
        for d in range(self.D):
            self._disarm(e, d)
        self._episode_start(e)

    def reset(self, mask=None):
        for e in range(self.E):
            if mask is None or mask[e]:
                self._reset_env(e)
        return self._observe(after_reset=np.ones(self.E, bool) if mask is None else np.asarray(mask, bool))

    # -------------------------------------------------------------- navigators
    def _nearest(self, e, src_pos, cand_slots):
        """argmin over the snapshot distance matrix, first index wins ties."""
        best, best_d = -1, np.inf
        for d in cand_slots:
            if not self.off_armed[e, d]:
                continue
            dist = np.linalg.norm(src_pos - self.off_pos[e, d])
            if dist < best_d:
                best, best_d = d, dist
        return best

    @staticmethod
    def _inside_cone(point, apex, base, degrees):
        """GeometryUtils.is_point_inside_cone (geometry_utils.py:6-29)."""
        ab = base - apex; ap = point - apex
        if np.linalg.norm(ap) > np.linalg.norm(ab):
            return False
        with np.errstate(invalid="ignore", divide="ignore"):
            cosang = np.dot(ap, ab) / (np.linalg.norm(ap) * np.linalg.norm(ab))
            ang = np.degrees(np.arccos(cosang))
        return bool(ang <= degrees / 2)

    def _path_clear(self, e, d, degrees):
        c = self.cfg
        if c.lm_nav == "air":
            return False
        b = np.asarray(c.building, dtype=np.float64)
        me = self.imu["position"][e, d]
        return not any(self._inside_cone(self.off_pos[e, j], me, b, degrees)
                       for j in range(c.n_lw) if self.off_armed[e, j])

    @staticmethod
    def _toward(target, me, speed):
        v = target - me
        n = np.linalg.norm(v)
        direction = v / n if n > 0 else v
        return np.array([*direction, speed])

    def _drive(self, e, d, command):
        self.setpoint[e, d] = dy.command_to_setpoint(command)

    def _navigate(self, e):
        c = self.cfg
        lws = range(c.n_lw)
        lms = range(c.n_lw, self.D)
        b = np.asarray(c.building, dtype=np.float64)
        for d in lms:
            if not self.armed[e, d]:
                continue
            me = self.imu["position"][e, d]
            alive = bool(self.off_armed[e, :c.n_lw].any())
            state = self.nav[e, d]
            if state == NAV_WAIT:
                if self._path_clear(e, d, 60):
                    self.nav[e, d] = NAV_BUILDING
                elif alive:
                    self.nav[e, d] = NAV_WINGMAN
                cmd = np.array([0.0, 0.0, 0.0, c.lm_speed])
            elif state == NAV_WINGMAN:
                if not alive:
                    self.nav[e, d] = NAV_BUILDING
                j = self._nearest(e, self.off_pos[e, d], lws)
                target = self.off_pos[e, j] if j >= 0 else np.zeros(3)
                cmd = self._toward(target, me, c.lm_speed)
            else:
                if not self._path_clear(e, d, 45):
                    self.nav[e, d] = NAV_WINGMAN
                cmd = self._toward(b, me, c.lm_speed)
            self._drive(e, d, cmd)
        self._navigate_allies(e, [j for j in lws if self.armed[e, j]][1:])

    def _gun_step(self, e):
        """Gun.current_step (gun.py:44-47): the env step of the last AGENT_STEP_BROADCAST."""
        return self.step_count[e]

    def _navigate_allies(self, e, allies):
        """drive_loyalwingmen (exp02_vFinal_task.py:237-242): scripted wingmen."""
        c = self.cfg
        lms = range(c.n_lw, self.D)
        for j in allies:
            if c.ally_mode == "stop":
                self._drive(e, j, np.array([0.0, 0.0, 0.0, c.ally_stop_mag]))
                continue
            me = self.imu["position"][e, j]
            available = self.ammo[e, j] <= 0 or c.cooldown_steps <= self._gun_step(e) - self.last_fired[e, j]
            if available or self.ammo[e, j] <= 0:
                i = self._nearest(e, self.off_pos[e, j], lms)
                target = self.off_pos[e, i]
            else:
                target = self.formation[e, j]
            self._drive(e, j, self._toward(target, me, c.bt_speed))

    # ---------------------------------------------------------------- dynamics
    def _substeps(self):
        c, prm = self.cfg, self.prm
        for _ in range(c.substeps):
            act = self.armed
            s = dy.imu_state(self.pos, self.quat, self.vel, self.omega)
            s["quaternion"] = dy.quat_from_euler(s["attitude"])
            for k in self.imu:
                self.imu[k] = np.where(act[..., None], s[k], self.imu[k])
            pid = self.pid.copy()
            pwm = dy.control_update(pid, s, self.setpoint, 6, prm)
            self.pid = np.where(act[..., None], pid, self.pid)
            if prm.noise_ratio != 0.0:
                sub = np.broadcast_to(np.arange(self.D, dtype=np.uint32), (self.E, self.D))
                noise = px.normal4(self.seed, self.env_ids[:, None], self.phys_ctr[:, None].astype(np.uint32), sub)
            else:
                noise = np.zeros((self.E, self.D, 4))
            thr, f, t = dy.actuate(self.throttle, pwm, s["velocity"], noise, prm)
            p, q, v, w = dy.rigid_body_step(self.pos, self.quat, self.vel, self.omega, f, t, prm)
            a3 = act[..., None]
            self.throttle = np.where(a3, thr, self.throttle)
            self.pos = np.where(a3, p, self.pos); self.quat = np.where(a3, q, self.quat)
            self.vel = np.where(a3, v, self.vel); self.omega = np.where(a3, w, self.omega)
            self.phys_ctr += 1

    # ------------------------------------------------------------- engagement
    def _gun_available(self, e, j):
        """Gun.is_available (gun.py:56-75)."""
        if self.ammo[e, j] <= 0:
            return True
        return self.cfg.cooldown_steps <= self._gun_step(e) - self.last_fired[e, j]

    def _gun_state(self, e, j):
        """Gun.get_state (gun.py:101-113)."""
        c = self.cfg
        wait = max(c.cooldown_steps - (self._gun_step(e) - self.last_fired[e, j]), 0)
        mx = c.munition if c.munition > 0 else 1
        return np.array([self.ammo[e, j] / mx, wait / c.cooldown_steps, int(self._gun_available(e, j))])

    def _in_range(self, e, thr):
        """identify_invaders_in_range (offsets_handler.py:283-309): {lw: [lm sorted by d]}."""
        c = self.cfg
        out = {}
        for j in range(c.n_lw):
            if not self.off_armed[e, j]:
                continue
            lst = []
            for d in range(c.n_lw, self.D):
                if not self.off_armed[e, d]:
                    continue
                dist = np.linalg.norm(self.off_pos[e, j] - self.off_pos[e, d])
                self.min_margin[e] = min(self.min_margin[e], abs(dist - thr))
                if dist < thr:
                    lst.append((d, dist))
            lst.sort(key=lambda x: x[1])
            if lst:
                out[j] = [d for d, _ in lst]
        return out

    def _middle(self, e):
        """Task.on_step_middle (exp02_vFinal_task.py:284-318)."""
        c = self.cfg
        ev = {"shots": [], "explosions": [], "origin": []}
        self._offsets(e)
        if c.reward == "v2full":           # update_building_life (exp02_v2_full_task.py)
            cnt = sum(1 for d in range(c.n_lw, self.D) if self.off_armed[e, d]
                      and np.linalg.norm(self.off_pos[e, d]) < 0.2)
            self.building_life[e] = max(self.building_life[e] - cnt, 0)
        agent_shots = ally_shots = 0
        for j, targets in self._in_range(e, c.shoot_range).items():
            can_fire = self._gun_available(e, j) and self.ammo[e, j] > 0
            if not can_fire:
                continue
            self.ammo[e, j] -= 1
            self.last_fired[e, j] = self._gun_step(e)
            hit = self._hit_u(e) < c.fire_probability
            ev["shots"].append((j, targets[0], bool(hit)))
            if hit:
                self._disarm(e, targets[0])
                if j == 0: agent_shots += 1
                else: ally_shots += 1
        exploded = ally_suicide = agent_suicide = 0
        for j, targets in self._in_range(e, c.explosion_range).items():
            self._disarm(e, j); self._disarm(e, targets[0])
            ev["explosions"].append((j, targets[0]))
            if c.reward == "v2full":
                exploded += 1
            elif self.ammo[e, j] == 0 and j == 0: agent_suicide += 1
            elif self.ammo[e, j] == 0: ally_suicide += 1
            else: exploded += 1
        self.agent_kills[e] += agent_shots; self.allies_kills[e] += ally_shots
        self.deads[e] += exploded
        for d in range(c.n_lw, self.D):     # process_invaders_in_origin
            if self.off_armed[e, d]:
                n0 = np.linalg.norm(self.off_pos[e, d])
                self.min_margin[e] = min(self.min_margin[e], abs(n0 - 0.2))
                if n0 < 0.2:
                    self._disarm(e, d); ev["origin"].append(d)
        if c.reward == "v2full":
            reward = self._reward_v2full(e, agent_shots + ally_shots, exploded)
        else:
            reward = self._reward_vfinal(e, agent_shots, ally_shots, exploded, ally_suicide, agent_suicide)
        if agent_shots + ally_shots > 0:
            self.max_step[e] += c.step_increment
        done = self._termination(e)
        self.events.append((int(self.step_count[e]), e, ev))
        return reward, done

    def _outside_dome(self, e, slots):
        out = 0
        for d in slots:
            if self.off_armed[e, d]:
                n0 = np.linalg.norm(self.off_pos[e, d])
                self.min_margin[e] = min(self.min_margin[e], abs(n0 - self.cfg.dome_radius))
                out += n0 > self.cfg.dome_radius
        return out

    def _reward_vfinal(self, e, agent_shots, ally_shots, exploded, ally_suicide, agent_suicide):
        """exp02_vFinal_task.py:422-514."""
        c = self.cfg
        score = bonus = penalty = 0.0
        g = self._gun_state(e, 0)
        munition, reload_progress, gun_available = g
        position = self.imu["position"][e, 0]
        velocity = self.imu["velocity"][e, 0]
        distance_to_origin = float(np.linalg.norm(position))
        # identify_closest_ally / identify_closest_invader on the snapshot
        src = -1
        if self.off_armed[e, 0]:
            allies = [j for j in range(c.n_lw) if self.off_armed[e, j]]
            if len(allies) <= 1:
                src = 0
            else:
                bd, src = np.inf, -1
                for j in allies:
                    if j == 0: continue
                    dd = np.linalg.norm(self.off_pos[e, j] - self.off_pos[e, 0])
                    if dd < bd: bd, src = dd, j
        target = self._nearest(e, self.off_pos[e, src], range(c.n_lw, self.D)) if src >= 0 else -1
        target_position = self.imu["position"][e, target] if target > -1 else np.zeros(3)
        current = float(np.linalg.norm(position - target_position))
        self.reward_margin[e] = min(self.reward_margin[e], abs(self.last_closest[e] - current - 0.01))
        if 0.01 < self.last_closest[e] - current and (gun_available == 1 or munition == 0):
            bonus += c.vel_bonus * np.linalg.norm(velocity)
        self.last_closest[e] = current
        if gun_available == 1 or munition == 0:
            score = -current
        else:
            score = current * (2 * reload_progress - 1)
        if agent_shots > 0 or agent_suicide > 0:
            bonus += (agent_shots + agent_suicide) * 1000
        if ally_shots > 0 or ally_suicide > 0:
            bonus += 0.5 * (ally_shots + ally_suicide) * 1000
        elif exploded > 0:
            penalty += 1000 * exploded
        if position[2] < -5:
            penalty += (-5 - position[2]) / (-5 + 6) * 1000
        if self._outside_dome(e, range(c.n_lw)) > 0:
            penalty += 1000
        if distance_to_origin > c.born_radius - 2:
            penalty += distance_to_origin - c.born_radius - 2
        return score + bonus - penalty

    def _reward_v2full(self, e, shots, exploded):
        """exp02_v2_full_task.py compute_reward."""
        c = self.cfg
        bonus = penalty = 0.0
        kills = self.agent_kills[e] + self.allies_kills[e]
        position = self.imu["position"][e, 0]
        if shots > 0:
            bonus += (shots + kills / 10) * 1000
        if self.ammo[e, 0] == 0 and exploded > 0:
            bonus += (shots + kills / 10) * 1000
        elif exploded > 0:
            penalty += 1000 * exploded
        if position[2] < 0.01:
            penalty += 1000
        if self._outside_dome(e, range(c.n_lw)) > 0:
            penalty += 1000
        if self.building_life[e] < 1:
            penalty += 1000 * (1 - self.building_life[e])
        dist = float(np.linalg.norm(position))
        if dist > c.born_radius:
            penalty += dist - c.born_radius
        return 0 + bonus - penalty

    def _termination(self, e):
        """exp02_vFinal_task.py:516-568 / exp02_v2_full_task.py."""
        c = self.cfg
        if self._gun_step(e) > self.max_step[e]:      # Task.current_step, same broadcast as the guns
            return True
        if not self.armed[e, c.n_lw:].any() and self.round[e] >= c.n_lm:
            return True
        if c.reward == "v2full" and self.building_life[e] <= 0:
            return True
        if self._outside_dome(e, range(c.n_lw)) > 0:
            return True
        if self._outside_dome(e, range(c.n_lw, self.D)) > 0:
            return True
        if not self.armed[e, :c.n_lw].any():
            return True
        if not self.armed[e, 0]:
            return True
        z = self.imu["position"][e, 0, 2]
        zthr = 0.01 if c.reward == "v2full" else -5.99
        self.min_margin[e] = min(self.min_margin[e], abs(z - zthr))
        return bool(z < zthr)

    def _step_end(self, e):
        """Task.on_step_end (exp02_vFinal_task.py:320-332) + advance_round (:154-174)."""
        c = self.cfg
        lm_alive = self.armed[e, c.n_lw:].any()
        if not lm_alive and self.round[e] >= c.n_lm:
            return
        if not lm_alive and self.armed[e, :c.n_lw].any():
            self.round[e] += 1 if self.round[e] < c.n_lm else c.n_lm
            self._setup_round(e, int(self.round[e]))
            self._offsets(e)
            self.nav[e] = NAV_WAIT

    # ---------------------------------------------------------------- observe
    def _lidar(self, e, obs_slot):
        c = self.cfg
        if not self.armed[e, obs_slot]:
            # own publisher was removed from the ring (disarm -> terminate): update_data returns
            # early, the previous sphere stays and features = [] (fused_lidar.py:160-166)
            return self.lidar_obs[e].copy(), np.full((N_THETA, N_PHI), -1, dtype=np.int32)
        others = [d for d in range(self.D) if d != obs_slot and self.armed[e, d]]
        types = [LW_TYPE if d < c.n_lw else LM_TYPE for d in others]
        return lidar_project(self.imu["position"][e, obs_slot], self.imu["quaternion"][e, obs_slot],
                             self.imu["position"][e, others], types, others, c.lidar, 2 * c.dome_radius)

    def _observe(self, after_reset=None):
        c = self.cfg
        E = self.E
        inertial = np.zeros((E, 15), dtype=np.float32)
        max_speed = 1 * 10 * (1000 / 3600)      # quadcopter.py:589-600
        for e in range(E):
            if not (after_reset is not None and after_reset[e]):
                self.lidar_obs[e], self.lidar_ids[e] = self._lidar(e, 0)
            im = self.imu
            v = np.concatenate([
                np.clip(im["position"][e, 0] / c.dome_radius, -1, 1),
                np.clip(im["velocity"][e, 0] / max_speed, -1, 1),
                np.clip(im["attitude"][e, 0] / np.pi, -1, 1),
                np.clip(im["angular_rate"][e, 0] / (2 * np.pi), -1, 1),
                self._gun_state(e, 0)])
            inertial[e] = v.astype(np.float32)
        return {"lidar": self.lidar_obs.copy(), "inertial_data": inertial,
                "last_action": self.last_action.astype(np.float32)}

    # --------------------------------------------------------------------- step
    def step(self, actions):
        """Env.step (exp02_vFinal_environment.py:155-177) for every env."""
        actions = np.asarray(actions, dtype=np.float64)
        E = self.E
        self.last_action = actions.copy()
        for e in range(E):
            self._drive(e, 0, actions[e])
            self._navigate(e)
        self._substeps()
        self.step_count += 1
        reward = np.zeros(E); done = np.zeros(E, dtype=bool)
        for e in range(E):
            reward[e], done[e] = self._middle(e)
        info = {"agent_kills": self.agent_kills.copy(), "allies_kills": self.allies_kills.copy(),
                "deads": self.deads.copy(), "current_wave": self.round.copy(),
                "building_life": self.building_life.copy()}
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

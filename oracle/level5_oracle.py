"""CPU oracle -- batched restatement of the threatsense level5 "C1 fusion" env step.

TEST INFRASTRUCTURE, not the product (same rules as env_oracle.py).  Pinned against the reference's own code:
oracle/make_golden_level5.py executes ``Level5C1FusionEnvironment`` from /root/reference/src (through
oracle/refshim) with the same Philox streams and records trajectories under tests/golden/level5_*.npz;
tests/test_oracle_golden_level5.py replays them through this file.

What each block follows (paths under /root/reference/src):
  step / reset order ..... threatsense/level5/level5_envrionment.py:203-288 (Level5Environment.reset/step/advance_step)
  observation ............ threatsense/level5/level5_c1_fusion_environment.py:20-57 (armed wingmen update their LiDAR,
                           the agent's stacked spheres + validity mask + inertial/gun vector + the agent's last command)
  task ................... threatsense/level5/components/tasks_management/tasks/level5_c1_fusion_task.py
                           constants :78-111, waves :131-183, episode :253-283, step hooks :285-350, engagement :352-424,
                           reward :434-484 (clipped, one-shot ``last_distance``), termination :488-545
  agent choice ........... threatsense/level5/components/entities_manager.py:350-383 (a random wingman)
  own sphere + features .. core/entities/quadcopters/components/sensors/fused_lidar.py:143-217
  ring semantics ......... .../sensors/components/lidar_buffer.py:54-75,104-145,208-232,289-292,363-437;
                           .../sensors/interfaces/base_lidar.py:44-85; core/entities/quadcopters/quadcopter.py:259-330
  stack ................... fused_lidar.py:73-109 (bootstrap, _build_valid_spheres), :223-269 (read_data, randomize_stack),
                           :293-326 (_pad_sphere_stack)
  neighbour re-framing ... .../sensors/components/lidar_math.py:16-22,25-34,53-83,186-260,262-324
  step-broadcast id clash  core/notification_system/message_hub.py:39-65 + level5_envrionment.py:276-281 + gun.py:44-47:
                           the environment broadcasts the step as publisher 0, which is also the body id of the first
                           munition (spawned before the ground plane); every ``disarm`` of that munition re-broadcasts
                           {"termination": True} on the step topic and zeroes ``current_step`` of every gun and of
                           the task until the next broadcast.

The agent's ring is modelled by what it provably contains (derivation in DESIGN.md section 3b): within an episode a
wingman P is a publisher from the reset observation until it is disarmed, the snapshot created at ring step s holds
P's pose of step s+1 and -- for a neighbour -- the features P broadcast at step s (none at s = 0), for the observer
itself the features of step s+1 (FusedLIDAR.update_data overwrites slot 1 of its own ring).
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass

import numpy as np

from . import dynamics as dy
from . import philox as px
from .env_oracle import EnvOracle, LM_TYPE, LW_TYPE, N_PHI, N_THETA, NAV_WAIT, Stage03Config

N_STACK = 6          # n_neighbors_max + 1 (fused_lidar.py:59, :307)
RING = 10            # max_buffer_size (base_lidar.py:37)


@dataclass
class Level5Config(Stage03Config):
    n_lw: int = 2                         # NUM_PURSUERS (level5_c1_fusion_task.py:83)
    n_lm: int = 10                        # MAX_NUM_INVADERS
    munition: int = 49                    # (4 + 10) * 7 // 2
    initial_invaders: int = 4
    invaders_per_round: int = 1
    max_rounds: int = 7                   # ceil((10 - 4) / 1 + 1)
    lidar_radius: float = 40.0
    # Level5FusionEnvironment (level5_fusion_environment.py) = the BASE Level5Environment + Level5FusionTask:
    l5_reward: str = "c1"                 # "fusion": level5_fusion_task.py:448-555
    base_env: bool = False                # base-class compute_observation (level5_envrionment.py:296-334): every
                                          # wingman, armed or not, updates its LiDAR; called 3x per step and per reset
                                          # (observation + info["student_observation"] + info["teacher_observation"],
                                          # :262-263,291-292,336-346); last_action is the env's, zeroed by reset (:153-155)
    # Level5DumbMultiObs (level5_dumb_multiobs.py) + Level5DumbMultiObjectTask: the data-collection env
    agent_bt: bool = False                # the agent flies the behaviour tree too (dumb task :255-266); no RL action
    agent_death_ends: bool = True         # dumb task :602-606 has the agent-dead termination commented out
    multi_obs: bool = False               # compute_info (:112-150): every ARMED wingman updates its LiDAR and yields a
                                          # student observation + its last command as the teacher action
    # Level52BTEvaluationEnvironment (level5_eval_2bt_environment.py) + Level52BTEvaluationTask
    random_agent: bool = True             # False: Teacher_Student=False, no agent is chosen (agent_id stays -1, no draw)
    z_end: bool = True                    # False: compute_termination (:430-468) has no z < -5.99 test
    no_obs: bool = False                  # True: step/reset never update a LiDAR (observation {} and no fusion draws)


LEVEL5_C1 = Level5Config()
# level5_fusion_task.py:81-112: 6 wingmen, 5 -> 30 munitions in 6 waves of +5, (5 + 30) * 6 // 2 = 105 rounds each
LEVEL5_FUSION = Level5Config(n_lw=6, n_lm=30, munition=105, initial_invaders=5, invaders_per_round=5, max_rounds=6,
                             l5_reward="fusion", base_env=True)
# level5_dumb_multiobject_task.py:81-107: 7 wingmen, 5 -> 30 munitions in 26 waves of +1, (5 + 30) * 26 // 2 = 455 rounds
LEVEL5_DUMB = Level5Config(n_lw=7, n_lm=30, munition=455, initial_invaders=5, invaders_per_round=1, max_rounds=26,
                           l5_reward="fusion", agent_bt=True, agent_death_ends=False, multi_obs=True)
# level5_2bt_evaluation_task.py:82-113: 2 behaviour-tree wingmen, 5 -> 30 munitions (+1 per wave), 455 rounds each,
# MAX_STEP = 300 + 10 * 100 and never incremented, no reward; info = kills per drone, deads, current wave
LEVEL5_EVAL2BT = Level5Config(n_lw=2, n_lm=30, munition=455, initial_invaders=5, invaders_per_round=1, max_rounds=26,
                              max_step=1300, step_increment=0, l5_reward="none", agent_bt=True, agent_death_ends=False,
                              z_end=False, no_obs=True, random_agent=False)


def fused_features(own_pos, own_quat, ent_pos, ent_type, ent_id, radius=40.0):
    """FusedLIDAR.update_data: sphere, winner ids and the kept feature list (r_n, theta, phi, type, id) in the order
    the cells were first claimed (lidar_math.add_features keeps a dict keyed by cell)."""
    sphere = np.ones((3, N_THETA, N_PHI), dtype=np.float32)
    ids = np.full((N_THETA, N_PHI), -1, dtype=np.int32)
    own_p = np.asarray(own_pos).astype(np.float32)
    own_q = np.asarray(own_quat).astype(np.float32)
    nsq = np.dot(own_q, own_q)
    qinv = np.array([-own_q[0], -own_q[1], -own_q[2], own_q[3]], dtype=np.float32) / nsq
    Rinv = dy.rot_from_quat(qinv.astype(np.float64))
    kept = {}
    for p, etype, eid in zip(ent_pos, ent_type, ent_id):
        p = np.asarray(p).astype(np.float32)
        rel = Rinv @ (p.astype(np.float64) - own_p.astype(np.float64))
        rn, theta, phi = _spherical(rel, radius)
        ti, pj = _cell(theta, phi)
        if rn < sphere[0, ti, pj]:
            sphere[0, ti, pj] = rn; sphere[1, ti, pj] = etype / 5; sphere[2, ti, pj] = 0.1
            ids[ti, pj] = eid
            kept[(ti, pj)] = (float(rn), float(theta), float(phi), int(etype), int(eid))
    return sphere, ids, list(kept.values())


def _spherical(v, radius):
    x, y, z = v
    r = np.sqrt(x * x + y * y + z * z)
    if r == 0:
        return 0.0, 0.0, 0.0
    return float(np.clip(r / radius, 0, 1)), float(np.arccos(np.clip(z / r, -1.0, 1.0))), float(np.arctan2(y, x))


def _cell(theta, phi):
    ti = int(np.clip(int(theta / np.pi * N_THETA), 0, N_THETA - 1))
    pj = int(np.clip(int((phi + np.pi) / (2 * np.pi) * N_PHI), 0, N_PHI - 1))
    return ti, pj


def neighbor_sphere(feats, n_pos, n_quat, o_pos, o_quat, own_id, age, radius=40.0):
    """LidarMath.neighbor_sphere_from_new_frame: denormalise the neighbour's kept features, neighbour frame -> world ->
    observer frame, rebin with the inverted priority (farther wins, anything beats an empty cell), time = age / 10."""
    sphere = np.ones((3, N_THETA, N_PHI), dtype=np.float32)
    n_pos = np.asarray(n_pos, dtype=np.float32).astype(np.float64)
    o_pos = np.asarray(o_pos, dtype=np.float32).astype(np.float64)
    Rn = dy.rot_from_quat(np.asarray(n_quat, dtype=np.float32).astype(np.float64))
    oq = np.asarray(o_quat, dtype=np.float32)
    qinv = np.array([-oq[0], -oq[1], -oq[2], oq[3]], dtype=np.float32) / np.dot(oq, oq)
    Rinv = dy.rot_from_quat(qinv.astype(np.float64))
    delta = float(np.clip(age / RING, 0.0, 1.0))
    for rn, theta, phi, etype, eid in feats:
        if eid == own_id:
            continue
        R = rn * radius
        cart = np.array([R * np.sin(theta) * np.cos(phi), R * np.sin(theta) * np.sin(phi), R * np.cos(theta)])
        rel = Rinv @ ((Rn @ cart + n_pos) - o_pos)
        rn2, th2, ph2 = _spherical(rel, radius)
        ti, pj = _cell(th2, ph2)
        cur = sphere[0, ti, pj]
        if (rn2 > cur) if cur < 1 else True:
            sphere[0, ti, pj] = rn2; sphere[1, ti, pj] = etype / 5; sphere[2, ti, pj] = delta
    return sphere


class Level5Oracle(EnvOracle):
    def __init__(self, cfg: Level5Config = LEVEL5_C1, n_envs: int = 1, seed: int = 0, env_offset: int = 0,
                 auto_reset: bool = False):
        E, L = n_envs, cfg.n_lw
        self.agent = np.zeros(E, dtype=np.int64)
        self.gun_step = np.zeros(E, dtype=np.int64)          # Gun.current_step == Task.current_step
        self.step_registered = np.zeros(E, dtype=bool)       # publisher 0 owns the step topic since the last broadcast
        self.last_distance = np.full(E, np.nan)              # ``hasattr(self, 'last_distance')`` one-shot
        self.obs_call = np.zeros(E, dtype=np.int64)
        self.last_cmd = np.zeros((E, 4))
        self.in_ring = np.zeros((E, L), dtype=bool)
        self.hist_pose = np.zeros((E, L, RING, 7), dtype=np.float32)
        self.hist_feat = [[[[] for _ in range(RING)] for _ in range(L)] for _ in range(E)]
        self.stack = np.ones((E, N_STACK, 3, N_THETA, N_PHI), dtype=np.float32)
        self.mask = np.zeros((E, N_STACK), dtype=bool)
        self.chosen = np.full((E, 4, 2), -1, dtype=np.int32)
        # base env only: the stack behind info["student_observation"] (second compute_observation call of a step)
        self.with_student = bool(cfg.base_env)
        self.student_stack = np.ones((E, N_STACK, 3, N_THETA, N_PHI), dtype=np.float32)
        self.student_mask = np.zeros((E, N_STACK), dtype=bool)
        self.student_chosen = np.full((E, 4, 2), -1, dtype=np.int32)
        # multi_obs: per wingman observer
        self.own_sphere = np.ones((E, L, 3, N_THETA, N_PHI), dtype=np.float32)
        self.last_cmd_lw = np.zeros((E, L, 4))               # Quadcopter.last_action of every wingman (survives resets)
        self.mo_stack = np.ones((E, L, N_STACK, 3, N_THETA, N_PHI), dtype=np.float32)
        self.mo_mask = np.zeros((E, L, N_STACK), dtype=bool)
        self.mo_chosen = np.full((E, L, 4, 2), -1, dtype=np.int32)
        super().__init__(cfg, n_envs, seed=seed, env_offset=env_offset, auto_reset=auto_reset)

    # ------------------------------------------------------------------ quirk
    def _gun_step(self, e):
        return self.gun_step[e]

    def _disarm(self, e, d):
        super()._disarm(e, d)
        if d < self.cfg.n_lw:
            self.in_ring[e, d] = False                        # messageHub.terminate -> close_buffer in every ring
        if d == self.cfg.n_lw and self.step_registered[e]:    # body id 0 == ENVIRONMENT_ID
            self.gun_step[e] = 0
            self.step_registered[e] = False

    def _broadcast(self, e):
        self.gun_step[e] = self.step_count[e]
        self.step_registered[e] = True

    # ----------------------------------------------------------------- episodes
    def _n_active(self, k):
        c = self.cfg
        return min((k - 1) * c.invaders_per_round + c.initial_invaders, c.n_lm)

    def _env_init(self, e):
        """Task.on_env_init + on_episode_start in Env.__init__ (level5_envrionment.py:176-184, task :253-273,609-660)."""
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
        self.agent[e] = int(self._spawn_u(e, 1)[0] * c.n_lw) if c.random_agent else 0   # set_agent(): a random wingman
        self._episode_start(e)

    def _setup_round(self, e, k):
        c = self.cfg
        for i in range(c.n_lm):
            self._disarm(e, c.n_lw + i)
        n = self._n_active(k)
        positions = self.generate_positions(e, n, c.born_radius)
        for i in range(n):
            self._replace(e, c.n_lw + i, positions[i])
            self._arm(e, c.n_lw + i)

    def _reset_env(self, e):
        """Level5Environment.reset (level5_envrionment.py:203-231) -> task.on_reset (:275-283); the ring is wiped by the
        step-0 broadcast (base_lidar.py:67-72), the agent's last command and ``last_distance`` survive."""
        c = self.cfg
        self.step_count[e] = 0
        self.max_step[e] = c.max_step
        self.agent_kills[e] = self.allies_kills[e] = self.deads[e] = 0
        self.last_closest[e] = c.dome_radius
        for d in range(self.D):
            self._disarm(e, d)
        self._episode_start(e)
        self._broadcast(e)

    # -------------------------------------------------------------- navigators
    def _navigate(self, e):
        super()._navigate(e)              # munitions; the ally list it passes on is replaced below

    def _navigate_allies(self, e, allies):
        """drive_loyalwingmen (level5_c1_fusion_task.py:244-248): every armed wingman except the agent."""
        c = self.cfg
        allies = [j for j in range(c.n_lw) if (c.agent_bt or j != self.agent[e]) and self.armed[e, j]]
        super()._navigate_allies(e, allies)

    def _drive(self, e, d, command):
        super()._drive(e, d, command)
        if d < self.cfg.n_lw:
            self.last_cmd_lw[e, d] = command          # Quadcopter.drive keeps the command (quadcopter.py:415-419)

    # ------------------------------------------------------------- engagement
    def _middle(self, e):
        """Task.on_step_middle (level5_c1_fusion_task.py:298-336)."""
        c = self.cfg
        ag = int(self.agent[e])
        ev = {"shots": [], "explosions": [], "origin": []}
        self._offsets(e)
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
                if j == ag: agent_shots += 1
                else: ally_shots += 1
        exploded = ally_suicide = agent_suicide = 0
        for j, targets in self._in_range(e, c.explosion_range).items():
            self._disarm(e, j); self._disarm(e, targets[0])
            ev["explosions"].append((j, targets[0]))
            if self.ammo[e, j] == 0 and j == ag: agent_suicide += 1
            elif self.ammo[e, j] == 0: ally_suicide += 1
            else: exploded += 1
        self.agent_kills[e] += agent_shots; self.allies_kills[e] += ally_shots
        self.deads[e] += exploded
        for d in range(c.n_lw, self.D):
            if self.off_armed[e, d]:
                n0 = np.linalg.norm(self.off_pos[e, d])
                self.min_margin[e] = min(self.min_margin[e], abs(n0 - 0.2))
                if n0 < 0.2:
                    self._disarm(e, d); ev["origin"].append(d)
        if c.l5_reward == "none":                         # "EVALUATION TASK DO NOT USES REWARD" (2bt task :418-427)
            reward = 0.0
        elif c.l5_reward == "fusion":
            reward = self._reward_fusion(e, agent_shots, ally_shots, exploded, ally_suicide, agent_suicide)
        else:
            reward = self._reward_c1(e, agent_shots, agent_suicide)
        if agent_shots + ally_shots > 0:
            self.max_step[e] += c.step_increment
        done = self._termination(e)
        self.events.append((int(self.step_count[e]), e, ev))
        return reward, done

    def _reward_c1(self, e, agent_shots, agent_suicide):
        """level5_c1_fusion_task.py:434-484."""
        c = self.cfg
        ag = int(self.agent[e])
        position = self.imu["position"][e, ag]
        target = self._nearest(e, self.off_pos[e, ag], range(c.n_lw, self.D)) if self.off_armed[e, ag] else -1
        target_position = self.imu["position"][e, target] if target > -1 else np.zeros(3)
        distance = float(np.linalg.norm(position - target_position))
        if np.isnan(self.last_distance[e]):
            self.last_distance[e] = distance
        reward = 0.0
        self.reward_margin[e] = min(self.reward_margin[e], abs(distance - self.last_distance[e]))
        if distance < self.last_distance[e]:
            reward += 10 * np.linalg.norm(self.imu["velocity"][e, ag])
        if agent_shots > 0:
            reward += 1.0 * agent_shots * 1000
        if agent_suicide > 0:
            reward -= 2.0 * agent_suicide * 1000
        return float(np.clip(reward, -3000.0, 3000.0))

    def _reward_fusion(self, e, agent_shots, ally_shots, exploded, ally_suicide, agent_suicide):
        """level5_fusion_task.py:448-555 (allies_dead is never passed by on_step_middle :325-331, so it is 0)."""
        c = self.cfg
        ag = int(self.agent[e])
        score = bonus = penalty = 0.0
        munition, _reload, gun_available = self._gun_state(e, ag)
        position = self.imu["position"][e, ag]
        distance_to_origin = float(np.linalg.norm(position))
        # identify_closest_ally / identify_closest_invader on the offsets snapshot (offsets_handler.py:167-281)
        src = -1
        if self.off_armed[e, ag]:
            allies = [j for j in range(c.n_lw) if self.off_armed[e, j]]
            if len(allies) <= 1:
                src = ag
            else:
                bd = np.inf
                for j in allies:
                    if j == ag: continue
                    dd = np.linalg.norm(self.off_pos[e, j] - self.off_pos[e, ag])
                    if dd < bd: bd, src = dd, j
        target = self._nearest(e, self.off_pos[e, src], range(c.n_lw, self.D)) if src >= 0 else -1
        target_position = self.imu["position"][e, target] if target > -1 else np.zeros(3)
        current = float(np.linalg.norm(position - target_position))
        MAX_REWARD, SAFE_RADIUS = 1000.0, 5.0
        if gun_available == 1 or munition == 0:
            score = -current
        else:
            score = +current
            if current < SAFE_RADIUS:
                penalty += ((SAFE_RADIUS - current) / max(SAFE_RADIUS, 1e-6)) * (0.50 * MAX_REWARD)
        self.reward_margin[e] = min(self.reward_margin[e], abs(current - self.last_closest[e] - 0.01))
        if (gun_available == 0 and munition > 0) and ((current - self.last_closest[e]) > 0.01):
            bonus += 0.10 * MAX_REWARD
        if agent_shots > 0:
            bonus += 1.0 * agent_shots * MAX_REWARD
        if ally_shots > 0 or ally_suicide > 0:
            bonus += 0.5 * (ally_shots + ally_suicide) * MAX_REWARD
        if agent_suicide > 0:
            penalty += 2.0 * agent_suicide * MAX_REWARD
        if exploded > 0:
            penalty += MAX_REWARD * exploded
        if position[2] < -5.0:
            penalty += min((-5.0 - position[2]) / 1.0, 1.0) * MAX_REWARD
        if self._outside_dome(e, range(c.n_lw)) > 0:
            penalty += MAX_REWARD
        if distance_to_origin > c.born_radius - 2:
            penalty += min((distance_to_origin - (c.born_radius - 2)) * 1.0, MAX_REWARD)
        self.last_closest[e] = current
        return float(np.clip(score + bonus - penalty, -3.0 * MAX_REWARD, 3.0 * MAX_REWARD))

    def _termination(self, e):
        """level5_c1_fusion_task.py:488-545."""
        c = self.cfg
        ag = int(self.agent[e])
        if self._gun_step(e) > self.max_step[e]:
            return True
        if not self.armed[e, c.n_lw:].any() and self.round[e] >= c.max_rounds:
            return True
        if self._outside_dome(e, range(c.n_lw)) > 0:
            return True
        if self._outside_dome(e, range(c.n_lw, self.D)) > 0:
            return True
        if not self.armed[e, :c.n_lw].any():
            return True
        if c.agent_death_ends and not self.armed[e, ag]:
            return True
        if not c.z_end:
            return False
        z = self.imu["position"][e, ag, 2]
        self.min_margin[e] = min(self.min_margin[e], abs(z + 5.99))
        return bool(z < -5.99)

    def _step_end(self, e):
        """Task.on_step_end (:338-350) + advance_round (:139-152)."""
        c = self.cfg
        lm_alive = self.armed[e, c.n_lw:].any()
        if not lm_alive and self.round[e] >= c.max_rounds:
            return
        if not lm_alive and self.armed[e, :c.n_lw].any():
            self.round[e] += 1 if self.round[e] < c.max_rounds else c.max_rounds
            self._setup_round(e, int(self.round[e]))
            self._offsets(e)
            self.nav[e] = NAV_WAIT

    # ---------------------------------------------------------------- observe
    def _fuse_u(self, e, sub, local):
        return float(px.uniform(self.seed, self.env_ids[e], px.STREAM_FUSE,
                                np.uint32(16 * int(self.obs_call[e]) + local), sub=sub))

    def _update_lidars(self, e, after_reset):
        """compute_observation's loop over the armed wingmen (update_data + read_data + broadcast)."""
        c = self.cfg
        cur = int(self.step_count[e])
        ag = int(self.agent[e])
        if after_reset:                                   # ring wiped; the broadcast re-registers the armed wingmen
            for P in range(c.n_lw):
                self.in_ring[e, P] = bool(self.armed[e, P])
                self.hist_feat[e][P][0] = []
            if self.armed[e, ag]:
                self.stack[e] = 1.0; self.mask[e] = False; self.chosen[e] = -1
                self.student_stack[e] = 1.0; self.student_mask[e] = False; self.student_chosen[e] = -1
            self.mo_stack[e] = 1.0; self.mo_mask[e] = False; self.mo_chosen[e] = -1
            return
        if c.base_env:
            # every wingman runs update_lidar in slot order; a disarmed one has no own snapshot any more (features [],
            # empty stack) but its LiDAR broadcast re-opens its buffer in the other rings: from then on it is a
            # candidate of get_random_neighborhood whose snapshots carry no pose -> never a sphere.  At the agent's
            # read_data of THIS call the wingmen in front of it have already broadcast.
            for P in range(ag):
                self.in_ring[e, P] = True
        for P in range(c.n_lw):
            if not self.armed[e, P]:
                continue
            others = [d for d in range(self.D) if d != P and self.armed[e, d]]
            types = [LW_TYPE if d < c.n_lw else LM_TYPE for d in others]
            sph, ids, feats = fused_features(self.imu["position"][e, P], self.imu["quaternion"][e, P],
                                             self.imu["position"][e, others], types, others, c.lidar_radius)
            self.hist_pose[e, P, cur % RING, :3] = self.imu["position"][e, P].astype(np.float32)
            self.hist_pose[e, P, cur % RING, 3:] = self.imu["quaternion"][e, P].astype(np.float32)
            self.hist_feat[e][P][cur % RING] = feats
            self.own_sphere[e, P] = sph
            if P == ag:
                self.lidar_obs[e], self.lidar_ids[e] = sph, ids
        if c.multi_obs:                                   # one read_data per armed wingman, draws keyed by its slot
            for P in range(c.n_lw):
                if self.armed[e, P]:
                    self.mo_stack[e, P], self.mo_mask[e, P], self.mo_chosen[e, P] = self._read_stack(e, 0, observer=P)
            return
        if not self.armed[e, ag]:
            if c.base_env:                                # read_data of the dead agent: [] padded to six empty spheres
                self.stack[e] = 1.0; self.mask[e] = False; self.chosen[e] = -1
                self.student_stack[e] = 1.0; self.student_mask[e] = False; self.student_chosen[e] = -1
                self.in_ring[e, :] = True
            return                                        # C1: the agent's flight state keeps the previous stack
        self.stack[e], self.mask[e], self.chosen[e] = self._read_stack(e, 0)
        if c.base_env:
            self.in_ring[e, :] = True                     # everybody has broadcast by the end of the call
            if self.with_student:
                # info["student_observation"] (level5_envrionment.py:291-292,342-346): compute_observation once more --
                # the ring updates are idempotent, every wingman (dead ones included) has an open buffer by now, and the
                # fusion draws are those of obs_call + 1
                self.student_stack[e], self.student_mask[e], self.student_chosen[e] = self._read_stack(e, 1)

    def _read_stack(self, e, call_offset, observer=None):
        """FusedLIDAR.read_data of an armed wingman (default: the agent) with the draws of compute_observation call
        obs_call + call_offset; the FUSE stream is keyed by the observer's slot."""
        c = self.cfg
        cur = int(self.step_count[e])
        ag = int(self.agent[e]) if observer is None else int(observer)
        u = lambda local: float(px.uniform(self.seed, self.env_ids[e], px.STREAM_FUSE,
                                           np.uint32(16 * (int(self.obs_call[e]) + call_offset) + local), sub=ag))
        spheres = [self.own_sphere[e, ag].copy()]
        n = 1 + int(u(0) * 4)
        cands = [P for P in range(c.n_lw) if self.in_ring[e, P]]
        k = min(n, len(cands))
        for i in range(k):
            j = i + int(u(1 + i) * (len(cands) - i))
            cands[i], cands[j] = cands[j], cands[i]
        chosen = np.full((4, 2), -1, dtype=np.int32)
        own_pose = self.hist_pose[e, ag, cur % RING]
        for i in range(k):
            P = cands[i]
            a = 1 + int(u(5 + i) * 9)
            chosen[i] = (P, a)
            s = cur - a
            if s < 0 or not self.armed[e, P]:             # (base env) re-registered after its death: no pose, no sphere
                continue
            pose = self.hist_pose[e, P, (s + 1) % RING]
            feats = self.hist_feat[e][ag][(s + 1) % RING] if P == ag else self.hist_feat[e][P][s % RING]
            spheres.append(neighbor_sphere(feats, pose[:3], pose[3:], own_pose[:3], own_pose[3:], ag, a, c.lidar_radius))
        order = list(range(N_STACK))
        for kk, i in enumerate(range(N_STACK - 1, 0, -1)):
            j = int(u(9 + kk) * (i + 1))
            order[i], order[j] = order[j], order[i]
        stack = np.ones((N_STACK, 3, N_THETA, N_PHI), dtype=np.float32)
        mask = np.zeros(N_STACK, dtype=bool)
        for dst, src in enumerate(order):
            if src < len(spheres):
                stack[dst] = spheres[src]; mask[dst] = True
        return stack, mask, chosen

    def _observe(self, after_reset=None):
        c = self.cfg
        E = self.E
        inertial = np.zeros((E, 15), dtype=np.float32)
        max_speed = 1 * 10 * (1000 / 3600)
        for e in range(E):
            # one compute_observation call of the reference per (env, obs): step obs for everyone, or the reset obs of
            # the envs that were just reset
            if (after_reset is None or after_reset[e]) and not c.no_obs:
                self._update_lidars(e, after_reset is not None)
                self.obs_call[e] += 3 if c.base_env else 1    # the 2nd and 3rd call only fill info
                if c.base_env and after_reset is not None:
                    self.last_cmd[e] = 0.0                    # init_globals: self.last_action = np.zeros(4)
            ag = int(self.agent[e])
            im = self.imu
            v = np.concatenate([
                np.clip(im["position"][e, ag] / c.dome_radius, -1, 1),
                np.clip(im["velocity"][e, ag] / max_speed, -1, 1),
                np.clip(im["attitude"][e, ag] / np.pi, -1, 1),
                np.clip(im["angular_rate"][e, ag] / (2 * np.pi), -1, 1),
                self._gun_state(e, ag)])
            inertial[e] = v.astype(np.float32)
        if c.multi_obs:
            # info["student_observations"] / info["teacher_actions"] of the armed wingmen (level5_dumb_multiobs.py:112-150)
            inertial_lw = np.zeros((E, c.n_lw, 15), dtype=np.float32)
            for e in range(E):
                for P in range(c.n_lw):
                    im = self.imu
                    inertial_lw[e, P] = np.concatenate([
                        np.clip(im["position"][e, P] / c.dome_radius, -1, 1), np.clip(im["velocity"][e, P] / max_speed, -1, 1),
                        np.clip(im["attitude"][e, P] / np.pi, -1, 1), np.clip(im["angular_rate"][e, P] / (2 * np.pi), -1, 1),
                        self._gun_state(e, P)]).astype(np.float32)
            return {"present": self.armed[:, :c.n_lw].copy(), "stacked_spheres": self.mo_stack.copy(),
                    "validity_mask": self.mo_mask.copy(), "inertial_data": inertial_lw,
                    "last_action": self.last_cmd_lw.astype(np.float32)}
        obs = {"stacked_spheres": self.stack.copy(), "validity_mask": self.mask.copy(), "inertial_data": inertial,
               "last_action": self.last_cmd.astype(np.float32), "lidar": self.lidar_obs.copy()}
        if self.with_student:
            obs["student_stacked_spheres"] = self.student_stack.copy()
            obs["student_validity_mask"] = self.student_mask.copy()
        return obs

    # --------------------------------------------------------------------- step
    def step(self, actions):
        """Level5Environment.step (level5_envrionment.py:236-266) for every env."""
        actions = np.asarray(actions, dtype=np.float64)
        E = self.E
        self.last_action = actions.copy()
        self.last_cmd = actions.copy()
        for e in range(E):
            if not self.cfg.agent_bt:
                self._drive(e, int(self.agent[e]), actions[e])
            self._navigate(e)
        self._substeps()
        self.step_count += 1
        reward = np.zeros(E); done = np.zeros(E, dtype=bool)
        for e in range(E):
            self._broadcast(e)
            reward[e], done[e] = self._middle(e)
        info = {"agent_kills": self.agent_kills.copy(), "allies_kills": self.allies_kills.copy(),
                "deads": self.deads.copy(), "current_wave": self.round.copy()}
        obs = self._observe()
        self.terminal_obs = obs
        for e in range(E):
            self._step_end(e)
        if self.auto_reset and done.any():
            obs = {k: v.copy() for k, v in obs.items()}
            for e in np.nonzero(done)[0]:
                self._reset_env(e)
            ar = np.zeros(E, bool); ar[done] = True
            new = self._observe(after_reset=ar)
            for k in obs:
                obs[k][done] = new[k][done]
        return obs, reward, done, info

"""Shared helpers for the parity tests."""
import dataclasses

import numpy as np

from oracle.env_oracle import EnvOracle, PRESETS as ORACLE_PRESETS


class Recording(dict):
    """A golden .npz read ONCE into memory.  (np.load is lazy: every ``rec["stacked"][k]`` on the NpzFile would
    decompress the whole array again -- tens of MB per access for the level5 recordings.)"""

    def __init__(self, path):
        with np.load(path) as z:
            super().__init__({k: z[k] for k in z.files})

    @property
    def files(self):
        return list(self.keys())


def load_recording(path):
    return Recording(path)


def oracle_cfg(name, **kw):
    return dataclasses.replace(ORACLE_PRESETS[name], **kw)


def kite_actions(orc: EnvOracle, rng: np.random.RandomState, chase_prob=0.9, ram=False):
    """Scripted pilot driven by the ORACLE's state (so both sides receive identical actions):
    close in on the nearest armed munition while the gun is ready, back off while it reloads."""
    c = orc.cfg
    E = orc.E
    a = np.zeros((E, 4))
    for e in range(E):
        lms = [d for d in range(c.n_lw, orc.D) if orc.armed[e, d]]
        if lms and rng.rand() < chase_prob:
            me = orc.imu["position"][e, 0]
            tgt = min(lms, key=lambda d: np.linalg.norm(orc.imu["position"][e, d] - me))
            v = orc.imu["position"][e, tgt] - me
            dist = max(np.linalg.norm(v), 1e-9)
            ready = orc._gun_available(e, 0) and orc.ammo[e, 0] > 0
            sign = 1.0 if (ready or ram or dist > 3.0) else -1.0
            a[e] = [*(sign * v / dist), rng.uniform(0.5, 1.0)]
        else:
            a[e] = [*rng.uniform(-1, 1, 3), rng.uniform(0, 1)]
    return a.astype(np.float32)


def oracle_state_dict(orc: EnvOracle):
    """Oracle state in the layout of BatchedThreatEngageEnv.set_state."""
    E = orc.E
    return {
        "pos": orc.pos.copy(), "quat": orc.quat.copy(), "vel": orc.vel.copy(), "omega": orc.omega.copy(),
        "throttle": orc.throttle.copy(), "pid": orc.pid.copy(), "armed": orc.armed.copy(),
        "off_armed": orc.off_armed.copy(), "nav": orc.nav.copy(), "last_fired": orc.last_fired.copy(),
        "ammo": orc.ammo.copy(), "imu_pos": orc.imu["position"].copy(), "formation": orc.formation.copy(),
        "step": orc.step_count.copy(), "max_step": orc.max_step.copy(), "round": orc.round.copy(),
        "agent_kills": orc.agent_kills.copy(), "allies_kills": orc.allies_kills.copy(), "deads": orc.deads.copy(),
        "building_life": orc.building_life.copy(), "hit_ctr": orc.hit_ctr.copy(), "spawn_ctr": orc.spawn_ctr.copy(),
        "phys_ctr": orc.phys_ctr.copy(), "last_closest": orc.last_closest.copy(),
        "lw_init": orc.lw_init_pos.copy(),
    }

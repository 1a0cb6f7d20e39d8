"""The oracle (oracle/env_oracle.py) replayed against trajectories recorded from the
reference's OWN env/task/navigator/gun/LiDAR code (oracle/make_golden.py).

Tolerances: the recordings and the oracle are both float64 on the same restated
dynamics, so everything is compared at 1e-9 (reward, positions) / exact (flags,
info counters, LiDAR hit ids); the float32 observation tensors at 1e-6.
"""
import dataclasses
import glob
import os

import numpy as np
import pytest

from oracle.env_oracle import EnvOracle, PRESETS
from tests.util import load_recording

CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "stage03_*.npz")))


def _check_obs(rec, k, obs, orc, tag):
    for name, key in (("lidar", "lidar"), ("inertial", "inertial_data"), ("last_action", "last_action")):
        d = np.abs(rec[name][k].astype(np.float64) - obs[key][0].astype(np.float64)).max()
        assert d <= 1e-6, f"{tag}: {name} differs by {d}"
    if not rec["was_reset"][k]:
        assert (rec["ids"][k] == orc.lidar_ids[0]).all(), f"{tag}: LiDAR hit ids differ"
    assert (rec["armed"][k] == orc.armed[0]).all(), f"{tag}: armed flags differ"
    assert np.abs(rec["pos"][k] - orc.pos[0]).max() <= 1e-9, f"{tag}: positions differ"


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_oracle_matches_reference_recording(path):
    rec = load_recording(path)
    preset = str(rec["preset"])
    seed, env_index, n_steps, _ = (int(v) for v in rec["meta"])
    cfg = dataclasses.replace(PRESETS[preset], noise_ratio=float(rec["noise_ratio"]))
    orc = EnvOracle(cfg, 1, seed=seed, env_offset=env_index)
    obs = orc.reset()
    k = 0
    _check_obs(rec, k, obs, orc, "reset"); k += 1
    for t in range(n_steps):
        obs, r, done, info = orc.step(rec["actions"][t][None])
        assert abs(r[0] - rec["reward"][t]) <= 1e-9, f"step {t}: reward {r[0]} vs {rec['reward'][t]}"
        assert bool(done[0]) == bool(rec["done"][t]), f"step {t}: terminated flag"
        kills = [int(info["agent_kills"][0]), int(info["allies_kills"][0])]
        if preset == "exp02_v2_full":          # that task reports one pooled "kills" counter
            kills = [kills[0] + kills[1], 0]
        got = kills + [int(info["deads"][0]), int(info["current_wave"][0])]
        assert got == [int(v) for v in rec["info"][t]], f"step {t}: info {got} vs {rec['info'][t]}"
        _check_obs(rec, k, obs, orc, f"step {t}"); k += 1
        if done[0]:
            obs = orc.reset()
            _check_obs(rec, k, obs, orc, f"reset after step {t}"); k += 1
    assert [int(orc.spawn_ctr[0]), int(orc.hit_ctr[0]), int(orc.phys_ctr[0])] == [int(v) for v in rec["counters"]]


def test_golden_cases_exist():
    assert len(CASES) >= 7

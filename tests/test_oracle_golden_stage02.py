"""Stage02 oracle (oracle/stage02_oracle.py) replayed against recordings of the reference's own
PyflytL3EnviromentV2 / L3Stage1 code (oracle/make_golden_stage02.py)."""
import dataclasses
import glob
import os

import numpy as np
import pytest

from oracle.stage02_oracle import STAGE02, Stage02Oracle
from tests.util import load_recording

CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "stage02_*.npz")))


def _check(rec, k, obs, orc, tag):
    for name, key in (("lidar", "lidar"), ("inertial", "inertial_data"), ("last_action", "last_action")):
        d = np.abs(rec[name][k].astype(np.float64) - obs[key][0].astype(np.float64)).max()
        assert d <= 1e-6, f"{tag}: {name} differs by {d}"
    if not rec["was_reset"][k]:
        assert (rec["ids"][k] == orc.lidar_ids[0]).all(), f"{tag}: LiDAR hit ids"
    assert (rec["armed"][k] == orc.armed[0]).all(), f"{tag}: armed flags"
    assert np.abs(rec["pos"][k] - orc.pos[0]).max() <= 1e-9, f"{tag}: positions"
    assert int(rec["ammo"][k]) == int(orc.ammo[0, 0]), f"{tag}: agent ammunition"


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_stage02_oracle_matches_reference_recording(path):
    rec = load_recording(path)
    seed, env_index, n_steps, _ = (int(v) for v in rec["meta"])
    orc = Stage02Oracle(dataclasses.replace(STAGE02, noise_ratio=float(rec["noise_ratio"])), 1, seed=seed, env_offset=env_index)
    obs = orc.reset()
    k = 0
    _check(rec, k, obs, orc, "reset"); k += 1
    for t in range(n_steps):
        obs, r, done, info = orc.step(rec["actions"][t][None])
        assert abs(r[0] - rec["reward"][t]) <= 1e-9, f"step {t}: reward"
        assert bool(done[0]) == bool(rec["done"][t]), f"step {t}: terminated"
        _check(rec, k, obs, orc, f"step {t}"); k += 1
        if done[0]:
            obs = orc.reset()
            _check(rec, k, obs, orc, f"reset after {t}"); k += 1
    assert [int(orc.spawn_ctr[0]), int(orc.hit_ctr[0]), int(orc.phys_ctr[0])] == [int(v) for v in rec["counters"]]


def test_stage02_golden_cases_exist():
    assert len(CASES) >= 3

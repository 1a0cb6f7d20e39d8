"""Stage01 oracle (oracle/stage01_oracle.py) replayed against recordings of the reference's own
PyflytL2EnviromentModifiedV2 code (oracle/make_golden_stage01.py)."""
import dataclasses
import glob
import os

import numpy as np
import pytest

from oracle.stage01_oracle import STAGE01, Stage01Oracle
from tests.util import load_recording

CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "stage01_*.npz")))


def _check(rec, k, obs, orc, tag):
    for name, key in (("lidar", "lidar"), ("inertial", "inertial_data"), ("last_action", "last_action")):
        d = np.abs(rec[name][k].astype(np.float64) - obs[key][0].astype(np.float64)).max()
        assert d <= 1e-6, f"{tag}: {name} differs by {d}"
    if not rec["was_reset"][k]:
        assert (rec["ids"][k] == orc.lidar_ids[0]).all(), f"{tag}: LiDAR hit ids"
    assert np.abs(rec["pos"][k] - orc.pos[0]).max() <= 1e-9, f"{tag}: positions"


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_stage01_oracle_matches_reference_recording(path):
    rec = load_recording(path)
    seed, env_index, n_steps, _ = (int(v) for v in rec["meta"])
    orc = Stage01Oracle(dataclasses.replace(STAGE01, noise_ratio=float(rec["noise_ratio"])), 1, seed=seed, env_offset=env_index)
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
    assert [int(orc.spawn_ctr[0]), int(orc.phys_ctr[0])] == [int(rec["counters"][0]), int(rec["counters"][2])]


def test_stage01_golden_cases_exist():
    assert len(CASES) >= 2

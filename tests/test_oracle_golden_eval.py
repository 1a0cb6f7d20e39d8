"""oracle/eval_oracle.py (level4 Evaluation_Task / EvaluationEnvironment and exp05) against the recordings of the
reference's OWN classes (tests/golden/l4eval_*.npz, l4exp05_*.npz; generator oracle/make_golden_eval.py): the env's
observation / reward / terminated, the per-wingman info rows, and -- per policy-driven wingman and step -- the observation
the task handed to its policy (sphere, inertial + gun vector, shared last_action) and the action that came back."""
import glob
import os

import numpy as np
import pytest

from oracle.eval_oracle import DrivenOracle
from oracle.make_golden_eval import case_config
from tests.util import load_recording


def replay(path, make, step_fn):
    rec = load_recording(path)
    seed, env_index, n_steps, _ = (int(v) for v in rec["meta"])
    kind, spec, extra = str(rec["kind"]), tuple(str(s) for s in rec["drivers"]), eval(str(rec["extra"]))
    cfg, _ = case_config(kind, spec, extra)
    import dataclasses
    cfg = dataclasses.replace(cfg, noise_ratio=float(rec["noise_ratio"]))
    return rec, kind, cfg, seed, env_index, n_steps


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "l4e*.npz"))),
                         ids=lambda p: os.path.basename(p)[:-4])
def test_driven_oracle_replays_the_reference(path):
    rec, kind, cfg, seed, env_index, n_steps = replay(path, None, None)
    orc = DrivenOracle(cfg, 1, seed=seed, env_offset=env_index, auto_reset=False, salts=list(rec["salts"]))
    obs = orc.reset()
    k = 0

    def check_obs(tag):
        assert np.array_equal(obs["lidar"][0] < 1, rec["lidar"][k] < 1), f"{tag}: marked cells of the env observation"
        assert np.abs(obs["lidar"][0] - rec["lidar"][k]).max() < 1e-6, f"{tag}: sphere"
        assert np.abs(obs["inertial_data"][0] - rec["inertial"][k]).max() < 1e-6, f"{tag}: inertial"
        assert np.abs(obs["last_action"][0] - rec["last_action"][k]).max() < 1e-6, f"{tag}: last_action"
        assert np.array_equal(orc.armed[0], np.concatenate([rec["armed"][k][:cfg.n_lw], rec["armed"][k][cfg.n_lw:]])), f"{tag}: armed"
        assert np.abs(orc.pos[0] - rec["pos"][k])[orc.armed[0]].max(initial=0.0) < 1e-9, f"{tag}: positions"
    check_obs("reset"); k += 1
    n_calls = 0
    for t in range(n_steps):
        obs, r, d, info = orc.step(rec["actions"][t][None])
        tag = f"{os.path.basename(path)} step {t}"
        called = orc.nn_obs["called"][0]
        assert np.array_equal(called, rec["nn_called"][t]), f"{tag}: which wingmen were served by a policy"
        for j in np.nonzero(called)[0]:
            assert np.array_equal(orc.nn_obs["lidar"][0, j] < 1, rec["nn_lidar"][t, j] < 1), f"{tag}: wingman {j} policy sphere cells"
            assert np.abs(orc.nn_obs["lidar"][0, j] - rec["nn_lidar"][t, j]).max() < 1e-6, f"{tag}: wingman {j} policy sphere"
            assert np.abs(orc.nn_obs["inertial"][0, j] - rec["nn_inertial"][t, j]).max() < 1e-6, f"{tag}: wingman {j} policy inertial"
            assert np.abs(orc.nn_obs["last_action"][0, j] - rec["nn_last_action"][t, j]).max() < 1e-6, f"{tag}: wingman {j} shared last_action"
            assert np.abs(orc.nn_obs["action"][0, j] - rec["nn_action"][t, j]).max() < 1e-6, f"{tag}: wingman {j} action"
            n_calls += 1
        assert abs(r[0] - rec["reward"][t]) <= 1e-9 * max(1.0, abs(rec["reward"][t])), f"{tag}: reward"
        assert bool(d[0]) == bool(rec["done"][t]), f"{tag}: terminated"
        if kind == "evaluation":
            alive = rec["lw_alive"][t]
            assert np.array_equal(info["lw_alive"][0], alive), f"{tag}: armed wingmen in info"
            assert np.array_equal(info["lw_kills"][0][alive], rec["lw_kills"][t][alive]), f"{tag}: lw_kills"
            assert np.array_equal(info["lw_munitions"][0][alive], rec["lw_munitions"][t][alive]), f"{tag}: lw_munitions"
            if alive.any():
                assert int(info["current_wave"][0]) == int(rec["wave"][t]) and int(info["step"][0]) == int(rec["step"][t]), f"{tag}: wave/step"
        else:
            got = [int(info[key][0]) for key in ("agent_kills", "allies_kills", "deads", "current_wave")]
            assert got == [int(v) for v in rec["info4"][t]], f"{tag}: info"
        check_obs(tag); k += 1
        if d[0]:
            obs = orc.reset()
            check_obs(tag + " reset"); k += 1
    assert n_calls > 100
    assert np.array_equal([orc.spawn_ctr[0], orc.hit_ctr[0], orc.phys_ctr[0]], rec["counters"])

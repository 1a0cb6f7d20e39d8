"""Stage02 (level3 L3Stage1) CUDA path vs the oracle and vs the recordings of the reference's own code."""
import dataclasses
import glob
import os

import numpy as np
import pytest
import torch

from oracle.stage02_oracle import STAGE02, Stage02Oracle
from tests.util import kite_actions
from tests.util import load_recording

pytestmark = pytest.mark.gpu


def _make(E, seed, precision, auto_reset, noise=None, env_offset=0):
    from dronechase_b200 import BatchedThreatEngageEnv, preset
    kw = {} if noise is None else {"noise_ratio": noise}
    env = BatchedThreatEngageEnv(preset("stage02", **kw), n_envs=E, seed=seed, device=0, env_offset=env_offset,
                                 auto_reset=auto_reset, precision=precision, with_ids=True, with_terminal_obs=True)
    orc = Stage02Oracle(dataclasses.replace(STAGE02, **kw), E, seed=seed, env_offset=env_offset, auto_reset=auto_reset)
    return env, orc


def test_stage02_closed_loop_f64_exact_events():
    E, K = 48, 300
    env, orc = _make(E, 31, "f64", True)
    obs = env.reset(); ref = orc.reset()
    assert np.allclose(obs["inertial_data"].cpu().numpy(), ref["inertial_data"], atol=1e-6)
    rng = np.random.RandomState(2)
    ram = np.arange(E) % 3 == 0
    for t in range(K):
        a = kite_actions(orc, rng)
        a_ram = kite_actions(orc, np.random.RandomState(t), ram=True)
        a[ram] = a_ram[ram]
        obs, rew, done, info = env.step(torch.from_numpy(a).cuda())
        ref, r_ref, d_ref, i_ref = orc.step(a.astype(np.float64))
        assert np.array_equal(done.cpu().numpy().astype(bool), d_ref), f"step {t}: terminated"
        inf = env.info.cpu().numpy()
        assert np.array_equal(inf[:, 0], i_ref["agent_kills"]) and np.array_equal(inf[:, 2], i_ref["deads"]), f"step {t}: counters"
        assert np.allclose(rew.cpu().numpy(), r_ref, rtol=1e-6, atol=1e-5), f"step {t}: reward"
        assert np.allclose(obs["inertial_data"].cpu().numpy(), ref["inertial_data"], atol=1e-6), f"step {t}: inertial"
        assert np.array_equal(env.lidar_ids.cpu().numpy(), orc.lidar_ids), f"step {t}: LiDAR ids"
        assert np.allclose(obs["lidar"].cpu().numpy(), ref["lidar"], atol=1e-6), f"step {t}: sphere"
    st = env.get_state()
    assert np.array_equal(st["armed"], orc.armed) and np.array_equal(st["ammo"][:, :2], orc.ammo[:, :2])
    assert np.abs(st["pos"] - orc.pos).max() < 1e-7
    assert np.array_equal(st["spawn_ctr"], orc.spawn_ctr) and np.array_equal(st["hit_ctr"], orc.hit_ctr)
    assert orc.agent_kills.max() >= 1 or orc.hit_ctr.max() >= 1


def test_stage02_closed_loop_f32():
    E, K, MARGIN = 64, 200, 2e-4
    env, orc = _make(E, 33, "f32", True)
    env.reset(); orc.reset()
    rng = np.random.RandomState(5)
    excused = np.zeros(E, dtype=bool)
    for t in range(K):
        a = kite_actions(orc, rng)
        obs, rew, done, info = env.step(torch.from_numpy(a).cuda())
        orc.min_margin[:] = np.inf; orc.reward_margin[:] = np.inf
        ref, r_ref, d_ref, i_ref = orc.step(a.astype(np.float64))
        excused |= orc.min_margin < MARGIN
        ok = ~excused
        rok = ok & (orc.reward_margin > MARGIN)
        assert np.array_equal(done.cpu().numpy().astype(bool)[ok], d_ref[ok]), f"step {t}: terminated"
        assert np.array_equal(env.info.cpu().numpy()[ok, 0], i_ref["agent_kills"][ok]), f"step {t}: kills"
        assert np.allclose(rew.cpu().numpy()[rok], r_ref[rok], atol=5e-3, rtol=1e-5), f"step {t}: reward"
        assert np.allclose(obs["inertial_data"].cpu().numpy()[ok], ref["inertial_data"][ok], atol=5e-4), f"step {t}: inertial"
    assert excused.mean() < 0.1


def test_stage02_golden_replay_through_cuda(golden_dir):
    from dronechase_b200 import BatchedThreatEngageEnv, preset
    paths = sorted(glob.glob(os.path.join(golden_dir, "stage02_*.npz")))
    assert paths
    for path in paths:
        rec = load_recording(path)
        seed, env_index, n_steps, _ = (int(v) for v in rec["meta"])
        env = BatchedThreatEngageEnv(preset("stage02", noise_ratio=float(rec["noise_ratio"])), n_envs=1, seed=seed,
                                     env_offset=env_index, auto_reset=True, precision="f64", with_ids=True,
                                     with_terminal_obs=True)
        obs = env.reset()
        k = 1
        for t in range(n_steps):
            a = torch.from_numpy(rec["actions"][t][None].astype(np.float32)).cuda()
            obs, rew, done, info = env.step(a)
            assert abs(float(rew[0]) - rec["reward"][t]) <= 1e-3 + 1e-6 * abs(rec["reward"][t]), f"{path} step {t}: reward"
            assert bool(done[0]) == bool(rec["done"][t]), f"{path} step {t}: done"
            if not done[0]:
                assert np.abs(obs["lidar"].cpu().numpy()[0] - rec["lidar"][k]).max() < 1e-6, f"{path} step {t}: sphere"
                assert np.abs(obs["inertial_data"].cpu().numpy()[0] - rec["inertial"][k]).max() < 1e-6
                assert np.array_equal(env.lidar_ids.cpu().numpy()[0], rec["ids"][k]), f"{path} step {t}: ids"
                k += 1
            else:
                assert np.abs(env.terminal_obs["inertial_data"].cpu().numpy()[0] - rec["inertial"][k]).max() < 1e-6
                k += 1
                assert np.abs(obs["inertial_data"].cpu().numpy()[0] - rec["inertial"][k]).max() < 1e-6, "reset obs"
                assert np.abs(obs["lidar"].cpu().numpy()[0] - rec["lidar"][k]).max() < 1e-6, "sphere kept over reset"
                k += 1
        env.close()

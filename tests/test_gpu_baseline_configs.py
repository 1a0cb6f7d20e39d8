"""Every BASELINE.json config at its STATED size against the oracle (needs a B200).

  config 2  stage02, 1 agent (+ idle support wingman) vs 10 hovering munitions, 4,096 envs on one GPU
  config 5  swarm, 4 wingmen vs 64 munitions, 8,192 envs per GPU
  config 4' threatsense C1 (stacked 13x26 spheres), 65,536 envs -- the stacked-observation sibling of config 3
  (config 3 itself, 65,536 stage03 envs, is tests/test_gpu_full_size.py; config 1 is tests/test_gpu_stage01.py; the
  stand-alone LiDAR of config 4 at 16,384 envs is tests/test_gpu_lidar_full_size below.)

The oracle cannot run thousands of envs in seconds, but every random draw is a Philox function of (seed, GLOBAL env
index, stream, counter) and envs are independent: an oracle created with ``env_offset = w`` reproduces the envs
[w, w + W) of the big batch.  Windows at the start, the middle and the end of the batch are compared while all the other
envs fly random actions: float64 build -> exact events / counters / LiDAR ids, floats 1e-6; float32 PRODUCT build ->
closed loop with auto-reset under the margin protocol of test_gpu_stage03 (an env whose oracle reports a predicate
within MARGIN of its threshold is excused from exact comparison from then on)."""
import dataclasses

import numpy as np
import pytest
import torch

from oracle.env_oracle import EnvOracle
from oracle.level5_oracle import LEVEL5_C1, Level5Oracle
from oracle.stage02_oracle import STAGE02, Stage02Oracle
from tests.test_gpu_level5 import _kite as kite_level5
from tests.util import kite_actions, oracle_cfg

pytestmark = pytest.mark.gpu
STAGE02_10LM = dataclasses.replace(STAGE02, n_lm=10, initial_round=10)
INFO_COLS = ((0, "agent_kills"), (1, "allies_kills"), (2, "deads"), (3, "current_wave"))


def _random_actions(E, g):
    act = torch.rand(E, 4, device="cuda", generator=g)
    act[:, :3] = act[:, :3] * 2 - 1
    return act


def _window_run(env, orcs, windows, W, K, kite, precision, cols, tag0, margin=2e-4, ram_after=40):
    """Step the big batch K times; the windows get the scripted pilot computed on their oracle's state."""
    E = env.n_envs
    obs = env.reset()
    refs = [o.reset() for o in orcs]
    for w, ref in zip(windows, refs):
        assert np.allclose(obs["inertial_data"][w:w + W].cpu().numpy(), ref["inertial_data"], atol=1e-6), f"{tag0}: reset obs, window {w}"
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    rngs = [np.random.RandomState(10 + i) for i in range(len(windows))]
    excused = [np.zeros(W, dtype=bool) for _ in windows]
    seen = {"kills": 0, "episodes": 0, "wave": 0, "compared": 0}
    for t in range(K):
        act = _random_actions(E, g)
        win_a = []
        for o, rng, w in zip(orcs, rngs, windows):
            a = kite(o, rng, ram=(t > ram_after))
            act[w:w + W] = torch.from_numpy(a).cuda()
            win_a.append(a)
        obs, rew, done, info = env.step(act)
        for o, a, w, exc in zip(orcs, win_a, windows, excused):
            o.min_margin[:] = np.inf; o.reward_margin[:] = np.inf
            ref, r_ref, d_ref, i_ref = o.step(a.astype(np.float64))
            sl = slice(w, w + W)
            tag = f"{tag0} window {w} step {t}"
            if precision == "f32":
                exc |= o.min_margin < margin
            ok = ~exc
            rok = ok & (o.reward_margin > (1e-3 if precision == "f32" else 0.0)) if precision == "f32" else ok
            assert np.array_equal(done[sl].cpu().numpy().astype(bool)[ok], d_ref[ok]), f"{tag}: terminated"
            got = env.info[sl].cpu().numpy()
            for col, key in cols:
                assert np.array_equal(got[ok, col], i_ref[key][ok]), f"{tag}: {key}"
            ftol = dict(rtol=1e-6, atol=1e-5) if precision == "f64" else dict(rtol=1e-5, atol=5e-3)
            assert np.allclose(rew[sl].cpu().numpy()[rok], r_ref[rok], **ftol), f"{tag}: reward"
            assert np.allclose(obs["inertial_data"][sl].cpu().numpy()[ok], ref["inertial_data"][ok],
                               atol=1e-6 if precision == "f64" else 5e-4), f"{tag}: inertial"
            if precision == "f64" and env.lidar_ids is not None:
                assert np.array_equal(env.lidar_ids[sl].cpu().numpy(), o.lidar_ids), f"{tag}: LiDAR hit ids"
                assert np.allclose(obs["lidar"][sl].cpu().numpy(), ref["lidar"], atol=1e-6), f"{tag}: sphere"
            seen["kills"] = max(seen["kills"], int((i_ref["agent_kills"] + i_ref.get("allies_kills", 0)).max()))
            seen["wave"] = max(seen["wave"], int(np.max(i_ref.get("current_wave", 0))))
            seen["episodes"] += int(d_ref.sum())
            seen["compared"] += int(ok.sum())
    seen["excused"] = float(np.mean([e.mean() for e in excused]))
    return seen


# ------------------------------------------------------------------------------------------ config 2
@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_config2_stage02_10lm_4096_envs(precision):
    from dronechase_b200 import BatchedThreatEngageEnv, preset
    E, W, K, seed = 4096, 16, 170, 41
    windows = (0, 2040, E - W)
    env = BatchedThreatEngageEnv(preset("stage02_10lm"), n_envs=E, seed=seed, device=0, auto_reset=True, precision=precision,
                                 with_ids=True)
    orcs = [Stage02Oracle(STAGE02_10LM, W, seed=seed, env_offset=w, auto_reset=True) for w in windows]
    seen = _window_run(env, orcs, windows, W, K, kite_actions, precision, ((0, "agent_kills"), (2, "deads")),
                       f"stage02_10lm {precision}")
    assert seen["kills"] >= 1 and seen["episodes"] >= 1, f"scenario too tame: {seen}"
    assert seen["excused"] < 0.15 and seen["compared"] > 0.8 * len(windows) * W * K, seen
    env.close()


# ------------------------------------------------------------------------------------------ config 5
@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_config5_swarm_8192_envs(precision):
    """4 wingmen (agent + 3 on the behaviour tree) vs 64 munitions, all armed in the first wave; auto-reset on."""
    from dronechase_b200 import BatchedThreatEngageEnv, preset
    E, W, K, seed = 8192, 6, 110, 19
    windows = (0, 4093, E - W)
    env = BatchedThreatEngageEnv(preset("swarm"), n_envs=E, seed=seed, device=0, auto_reset=True, precision=precision,
                                 with_ids=True)
    orcs = [EnvOracle(oracle_cfg("swarm"), W, seed=seed, env_offset=w, auto_reset=True) for w in windows]
    seen = _window_run(env, orcs, windows, W, K, kite_actions, precision, INFO_COLS, f"swarm {precision}", ram_after=30)
    # 64 munitions close in on 4 wingmen at 0.4 m/s from r = 6: the first shots fall around step 75; an episode ends
    # when a munition gets within 0.2 m of the agent (explosion) -- both must have happened inside the windows
    assert seen["kills"] >= 1 and seen["episodes"] >= 1, f"scenario too tame: {seen}"
    # float32: 4 x 64 range predicates per env and step (ten times exp02's): more envs come within MARGIN of a threshold
    # during 110 steps and are excused from then on (39 % measured); the others must match exactly
    assert seen["excused"] < (0.05 if precision == "f64" else 0.55) and seen["compared"] > 0.45 * len(windows) * W * K, seen
    env.close()


def test_config5_swarm_wave_advance_f64():
    """A swarm env whose first wave is wiped out advances the round: shortened to 3 munitions per wave so that the wave
    logic of the D = 68 geometry (epw = 4) is reached in seconds -- exact against the oracle."""
    from dronechase_b200 import BatchedThreatEngageEnv, preset
    E, K, seed = 24, 260, 3
    cfg = preset("swarm", initial_round=2)
    env = BatchedThreatEngageEnv(cfg, n_envs=E, seed=seed, device=0, auto_reset=True, precision="f64", with_ids=True)
    orc = EnvOracle(oracle_cfg("swarm", initial_round=2), E, seed=seed, auto_reset=True)
    seen = _window_run(env, [orc], (0,), E, K, kite_actions, "f64", INFO_COLS, "swarm waves", ram_after=10**9)
    assert seen["kills"] >= 1 and seen["wave"] >= 3, f"no wave advance: {seen}"
    env.close()


# ------------------------------------------------------------------------------------------ threatsense C1 at 65,536 envs
def test_level5_c1_65536_envs_windows_f64():
    from dronechase_b200 import BatchedThreatEngageEnv, preset
    E, W, K, seed = 65536, 8, 120, 13
    windows = (0, 32764, E - W)
    env = BatchedThreatEngageEnv(preset("level5_c1"), n_envs=E, seed=seed, device=0, auto_reset=True, precision="f64")
    orcs = [Level5Oracle(LEVEL5_C1, W, seed=seed, env_offset=w, auto_reset=True) for w in windows]
    obs = env.reset()
    refs = [o.reset() for o in orcs]
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    rngs = [np.random.RandomState(20 + i) for i in range(len(windows))]
    kills = episodes = marked = 0
    for t in range(K):
        act = _random_actions(E, g)
        win_a = []
        for o, rng, w in zip(orcs, rngs, windows):
            a = kite_level5(o, rng, ram=(t > 50))
            act[w:w + W] = torch.from_numpy(a).cuda()
            win_a.append(a)
        obs, rew, done, info = env.step(act)
        for o, a, w in zip(orcs, win_a, windows):
            ref, r_ref, d_ref, i_ref = o.step(a.astype(np.float64))
            sl, tag = slice(w, w + W), f"level5_c1 window {w} step {t}"
            assert np.array_equal(done[sl].cpu().numpy().astype(bool), d_ref), f"{tag}: terminated"
            got = env.info[sl].cpu().numpy()
            for col, key in INFO_COLS:
                assert np.array_equal(got[:, col], i_ref[key]), f"{tag}: {key}"
            assert np.allclose(rew[sl].cpu().numpy(), r_ref, rtol=1e-6, atol=1e-5), f"{tag}: reward"
            assert np.allclose(obs["inertial_data"][sl].cpu().numpy(), ref["inertial_data"], atol=1e-6), f"{tag}: inertial"
            assert np.array_equal(obs["validity_mask"][sl].cpu().numpy(), ref["validity_mask"]), f"{tag}: validity mask"
            gs, ws = obs["stacked_spheres"][sl].cpu().numpy(), ref["stacked_spheres"]
            assert np.array_equal(gs < 1, ws < 1), f"{tag}: marked cells of the stack"
            assert np.abs(gs - ws).max() <= 1e-6, f"{tag}: stacked spheres"
            kills = max(kills, int(i_ref["agent_kills"].max())); episodes += int(d_ref.sum()); marked += int((ws < 1).sum())
    assert kills >= 1 and episodes >= 1 and marked > 5000, (kills, episodes, marked)
    env.close()


# ------------------------------------------------------------------------------------------ config 4
def test_config4_lidar_16384_envs_16_entities():
    """Stand-alone projection LiDAR at BASELINE's size (16,384 envs x 16 entities, 6 observers = 98,304 spheres): windows
    of envs against oracle/env_oracle.lidar_project (cells / ids exact, distances 5e-7), and the size-independent
    properties -- every marked cell holds an id, every id's distance is the minimum of the entities binned there."""
    from dronechase_b200 import lidar_project
    from oracle.env_oracle import lidar_project as oracle_project
    E, N, O = 16384, 16, 6
    rng = np.random.RandomState(4)
    v = rng.normal(size=(E, N, 3)); v /= np.linalg.norm(v, axis=-1, keepdims=True)
    pos = (v * rng.uniform(0, 1, (E, N, 1)) ** (1 / 3) * 6.0).astype(np.float32)
    q = rng.normal(size=(E, N, 4)); q /= np.linalg.norm(q, axis=-1, keepdims=True)
    quat = q.astype(np.float32)
    types = np.array([3] * O + [1] * (N - O), dtype=np.int32)
    alive = rng.rand(E, N) < 0.9
    alive[:, :O] = True
    sphere, ids = lidar_project(torch.from_numpy(pos).cuda(), torch.from_numpy(quat).cuda(), torch.from_numpy(types),
                                torch.from_numpy(alive), torch.arange(O, dtype=torch.int32), flavour="fused", radius=40.0,
                                with_ids=True)
    sphere, ids = sphere.cpu().numpy(), ids.cpu().numpy()
    assert sphere.shape == (E, O, 3, 13, 26) and ids.shape == (E, O, 13, 26)
    assert np.array_equal(sphere[:, :, 0] < 1, ids >= 0)                      # a marked cell names its entity
    assert np.all(sphere[:, :, 2][ids >= 0] == np.float32(0.1)) and np.all(sphere[:, :, 1][ids < 0] == 1)
    flag = sphere[:, :, 1][ids >= 0]
    want_flag = np.where(ids[ids >= 0] < O, np.float32(0.6), np.float32(0.2))
    assert np.array_equal(flag, want_flag)
    for e in list(range(0, 8)) + list(range(8190, 8196)) + list(range(E - 8, E)):
        for o in range(O):
            others = [d for d in range(N) if d != o and alive[e, d]]
            ref_s, ref_ids = oracle_project(pos[e, o], quat[e, o], pos[e, others], types[others], np.array(others),
                                            flavour="fused", radius=40.0)
            assert np.array_equal(ids[e, o], ref_ids), f"env {e} observer {o}: ids"
            assert np.abs(sphere[e, o] - ref_s).max() < 5e-7, f"env {e} observer {o}: sphere"

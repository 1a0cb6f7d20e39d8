"""Parity at BASELINE.json's full batch size (65,536 envs per GPU) -- needs a B200.

The oracle cannot run 65,536 envs in seconds, but envs are independent and every random draw is a Philox function of
(seed, GLOBAL env index, stream, counter): an oracle instantiated with ``env_offset = w`` for a window of envs
reproduces exactly those envs of the big batch.  Three windows are checked against the float64 build of the CUDA path --
the first envs, a window straddling the sub-batch border (32,768) and the last envs -- with exact events / counters /
LiDAR hit ids and 1e-6 floats, while the other 65,488 envs fly random actions.  Size-independent properties on the
float32 product build: the same batch run as 1 and as 2 sub-batches gives the same bits, and two runs give the same bits.
"""
import numpy as np
import pytest
import torch

from tests.util import kite_actions, oracle_cfg
from oracle.env_oracle import EnvOracle

pytestmark = pytest.mark.gpu

E_FULL = 65536
WINDOWS = (0, 32760, E_FULL - 16)
W = 16


@pytest.mark.parametrize("name", ["exp02_vFinal", "exp02_v2_full"])
def test_full_batch_windows_match_oracle_f64(name):
    from dronechase_b200 import BatchedThreatEngageEnv, preset
    seed, K = 77, 140
    env = BatchedThreatEngageEnv(preset(name), n_envs=E_FULL, seed=seed, device=0, auto_reset=True, precision="f64",
                                 with_ids=True, sub_batches=2)
    orcs = [EnvOracle(oracle_cfg(name), W, seed=seed, env_offset=w, auto_reset=True) for w in WINDOWS]
    obs = env.reset()
    refs = [o.reset() for o in orcs]
    for w, ref in zip(WINDOWS, refs):
        assert np.allclose(obs["inertial_data"][w:w + W].cpu().numpy(), ref["inertial_data"], atol=1e-6)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    rngs = [np.random.RandomState(10 + i) for i in range(len(WINDOWS))]
    kills = 0
    for t in range(K):
        act = torch.rand(E_FULL, 4, device="cuda", generator=g); act[:, :3] = act[:, :3] * 2 - 1
        win_a = []
        for o, rng, w in zip(orcs, rngs, WINDOWS):
            a = kite_actions(o, rng, ram=(t > 60))
            act[w:w + W] = torch.from_numpy(a).cuda()
            win_a.append(a)
        obs, rew, done, info = env.step(act)
        inf = env.info
        for o, a, w in zip(orcs, win_a, WINDOWS):
            ref, r_ref, d_ref, i_ref = o.step(a.astype(np.float64))
            sl = slice(w, w + W)
            tag = f"{name} window {w} step {t}"
            assert np.array_equal(done[sl].cpu().numpy().astype(bool), d_ref), f"{tag}: terminated"
            got = inf[sl].cpu().numpy()
            for col, key in ((0, "agent_kills"), (1, "allies_kills"), (2, "deads"), (3, "current_wave")):
                assert np.array_equal(got[:, col], i_ref[key]), f"{tag}: {key}"
            assert np.allclose(rew[sl].cpu().numpy(), r_ref, rtol=1e-6, atol=1e-5), f"{tag}: reward"
            assert np.allclose(obs["inertial_data"][sl].cpu().numpy(), ref["inertial_data"], atol=1e-6), f"{tag}: inertial"
            assert np.array_equal(env.lidar_ids[sl].cpu().numpy(), o.lidar_ids), f"{tag}: LiDAR hit ids"
            assert np.allclose(obs["lidar"][sl].cpu().numpy(), ref["lidar"], atol=1e-6), f"{tag}: sphere"
            kills = max(kills, int(i_ref["agent_kills"].max()))
    assert kills >= 1, "scenario too tame: no kill in the windows"
    env.close()


@pytest.mark.parametrize("name", ["exp02_vFinal", "level5_c1"])
def test_full_batch_invariants_f32(name):
    """Product build at full size: sub-batch split invariance and run-to-run determinism, through checksums of every output."""
    from dronechase_b200 import BatchedThreatEngageEnv

    def run(K_sub):
        env = BatchedThreatEngageEnv(name, n_envs=E_FULL, seed=5, device=0, sub_batches=K_sub)
        env.reset()
        g = torch.Generator(device="cuda"); g.manual_seed(3)
        sums = []
        for t in range(150):
            act = torch.rand(E_FULL, 4, device="cuda", generator=g); act[:, :3] = act[:, :3] * 2 - 1
            obs, rew, done, info = env.step(act)
            if t % 6 == 5:
                sums.append([float(rew.double().sum()), int(done.sum()), int(info.long().sum())]
                            + [float(v.double().sum()) for v in obs.values()])
        out = ({k: v.clone() for k, v in obs.items()}, rew.clone(), done.clone(), info.clone(), sums, env.stats.clone())
        env.close()
        return out
    a, b, c = run(1), run(2), run(2)
    for x, y, what in ((a, b, "1 vs 2 sub-batches"), (b, c, "run to run")):
        assert x[4] == y[4], f"{what}: checksums over time differ"
        for k in x[0]:
            assert torch.equal(x[0][k], y[0][k]), f"{what}: {k}"
        assert torch.equal(x[1], y[1]) and torch.equal(x[2], y[2]) and torch.equal(x[3], y[3]), what
    assert int(a[5][0]) > 10, "no episode ended: the invariants were not exercised through resets"

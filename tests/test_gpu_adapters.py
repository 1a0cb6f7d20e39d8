"""SB3 VecEnv contract and gymnasium facade over the batched simulator (needs a B200)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_vec_env_contract_and_autoreset():
    from dronechase_b200.vec_env import DroneChaseVecEnv
    n = 64
    venv = DroneChaseVecEnv("exp02_vFinal", n_envs=n, seed=3)
    assert venv.num_envs == n and venv.action_space.shape == (4,)
    obs = venv.reset()
    assert set(obs) == {"lidar", "inertial_data", "last_action"}
    assert obs["lidar"].shape == (n, 3, 13, 26) and obs["lidar"].dtype == np.float32
    assert obs["inertial_data"].shape == (n, 15) and obs["last_action"].shape == (n, 4)
    assert (obs["lidar"] == 1).all()                        # first reset: empty spheres
    rng = np.random.RandomState(0)
    finished = 0
    for t in range(330):
        a = np.concatenate([rng.uniform(-1, 1, (n, 3)), rng.uniform(0, 1, (n, 1))], axis=1).astype(np.float32)
        venv.step_async(a)
        obs, rew, dones, infos = venv.step_wait()
        assert rew.shape == (n,) and dones.shape == (n,) and dones.dtype == np.bool_ and len(infos) == n
        assert np.isfinite(rew).all() and (np.abs(obs["inertial_data"]) <= 1).all()
        assert ((obs["lidar"] >= 0) & (obs["lidar"] <= 1)).all()
        for i in np.nonzero(dones)[0]:
            info = infos[int(i)]
            finished += 1
            assert info["TimeLimit.truncated"] is False and "terminal_observation" in info
            assert info["terminal_observation"]["inertial_data"].shape == (15,)
            # SB3 semantics: the returned observation already belongs to the next episode
            assert np.allclose(obs["last_action"][i], 0) and np.allclose(obs["inertial_data"][i, 12:], [1, 0, 1])
        keep = ~dones
        assert np.allclose(obs["last_action"][keep], a[keep])
    assert finished >= 5                                    # explosions / time-outs happened and were auto-reset
    assert set(infos[0]) >= {"agent_kills", "allies_kills", "deads", "current_wave"}
    assert venv.env_is_wrapped(object) == [False] * n and venv.get_attr("num_envs")[0] == n
    venv.close()


def test_gym_facade_matches_batched_env():
    from dronechase_b200 import BatchedThreatEngageEnv
    from dronechase_b200.gym_env import Exp02vFinalEnvironment
    env = Exp02vFinalEnvironment(dome_radius=20, rl_frequency=15, GUI=False, seed=9)
    ref = BatchedThreatEngageEnv("exp02_vFinal", n_envs=1, seed=9, auto_reset=False)
    obs, info = env.reset()
    ref.reset()
    assert info == {} and obs["lidar"].shape == (3, 13, 26)
    for t in range(20):
        a = np.array([0.3, -0.2, 0.1, 0.8], dtype=np.float32)
        obs, r, term, trunc, info = env.step(a)
        o2, r2, d2, _ = ref.step(torch.from_numpy(a[None]).cuda())
        assert trunc is False and isinstance(r, float) and isinstance(term, bool)
        assert r == float(r2[0]) and np.array_equal(obs["inertial_data"], o2["inertial_data"][0].cpu().numpy())
    assert set(info) == {"agent_kills", "allies_kills", "deads", "current_wave"}
    env.close()
    with pytest.raises(ValueError):
        Exp02vFinalEnvironment(GUI=True)


def test_level5_vec_env_and_facade():
    from dronechase_b200.gym_env import Level5C1FusionEnvironment
    from dronechase_b200.vec_env import DroneChaseVecEnv
    n = 48
    venv = DroneChaseVecEnv("level5_c1", n_envs=n, seed=4)
    obs = venv.reset()
    assert set(obs) == {"stacked_spheres", "validity_mask", "inertial_data", "last_action"}
    assert obs["stacked_spheres"].shape == (n, 6, 3, 13, 26) and obs["validity_mask"].shape == (n, 6)
    assert obs["validity_mask"].dtype == np.bool_ and not obs["validity_mask"].any() and (obs["stacked_spheres"] == 1).all()
    assert venv.observation_space["stacked_spheres"].shape == (6, 3, 13, 26)
    rng = np.random.RandomState(1)
    finished = 0
    for t in range(320):
        a = np.concatenate([rng.uniform(-1, 1, (n, 3)), rng.uniform(0, 1, (n, 1))], axis=1).astype(np.float32)
        obs, rew, dones, infos = venv.step(a)
        valid = obs["validity_mask"]
        marked = (obs["stacked_spheres"][:, :, 0] < 1).any(axis=(2, 3))
        assert not (marked & ~valid).any(), "a padded sphere carries a hit"
        keep = ~dones
        assert (valid[keep].sum(1) >= 1).all() and (valid[keep].sum(1) <= 5).all()      # own + up to 4 neighbours
        assert not valid[dones].any()                                # reset observation: ring wiped
        assert np.allclose(obs["last_action"], a)                     # the agent's last command survives the reset
        finished += int(dones.sum())
        for i in np.nonzero(dones)[0]:
            assert "terminal_observation" in infos[int(i)]
    assert finished >= 3
    venv.close()
    env = Level5C1FusionEnvironment(GUI=False, seed=2)
    obs, info = env.reset()
    assert info == {} and obs["stacked_spheres"].shape == (6, 3, 13, 26) and obs["validity_mask"].shape == (6,)
    obs, r, term, trunc, info = env.step(np.array([0.1, 0.2, 0.3, 0.5], dtype=np.float32))
    assert info == {} and trunc is False and obs["validity_mask"].sum() >= 1
    env.close()
    from dronechase_b200.gym_env import Level5FusionEnvironment
    env = Level5FusionEnvironment(GUI=False, seed=2)
    obs, info = env.reset()
    assert obs["stacked_spheres"].shape == (6, 3, 13, 26) and obs["lidar"].shape == (2, 13, 26) and not obs["last_action"].any()
    assert set(info["teacher_observation"]) == {"lidar", "inertial_data", "last_action"}
    a = np.array([0.1, 0.2, 0.3, 0.5], dtype=np.float32)
    obs, r, term, trunc, info = env.step(a)
    assert np.allclose(obs["last_action"], a) and -3000.0 <= r <= 3000.0 and obs["validity_mask"].sum() >= 1
    so = info["student_observation"]           # level5_envrionment.py:342-346: a second, differently drawn stack
    assert set(so) == {"stacked_spheres", "validity_mask", "inertial_data", "last_action"}
    assert so["stacked_spheres"].shape == (6, 3, 13, 26) and so["validity_mask"].sum() >= 1
    assert np.array_equal(so["inertial_data"], obs["inertial_data"]) and np.array_equal(so["last_action"], obs["last_action"])
    assert ((so["stacked_spheres"] < 1).any(axis=(1, 2, 3)) <= so["validity_mask"]).all()
    env.close()
    from dronechase_b200.gym_env import Level5DumbMultiObs
    env = Level5DumbMultiObs(GUI=False, seed=4)
    obs, info = env.reset()
    assert obs.shape == (1,) and len(info["student_observations"]) == len(info["teacher_actions"]) == 7
    for _ in range(3):
        obs, r, term, trunc, info = env.step(np.zeros(4))
    assert obs.shape == (1,) and trunc is False and len(info["student_observations"]) == 7
    so = info["student_observations"][2]
    assert set(so) == {"stacked_spheres", "validity_mask", "inertial_data", "last_action"} and so["validity_mask"].any()
    assert np.array_equal(so["last_action"], info["teacher_actions"][2]) and abs(np.linalg.norm(so["last_action"][:3]) - 1) < 1e-5
    env.close()


@pytest.mark.parametrize("name", ["exp02_vFinal", "exp03_vFinal", "stage02", "level5_c1", "level5_fusion"])
def test_sparse_lidar_transfer_is_bit_identical(name):
    """The default adapter does not copy the sphere: the level4/3/2 families move the words that changed as an (index, value)
    list (dc_diff_hits + dc_host_apply_pairs) or, on request, mirror them into page-locked, device-mapped numpy arrays with a
    few PCIe writes per env (dc_mirror_hits); level5 moves the stack as a hit list and rebuilds it on the host
    (dc_host_scatter_stack; also available to the other families).  All must equal the dense copy."""
    from dronechase_b200.vec_env import DroneChaseVecEnv
    n = 96
    level5 = name.startswith("level5")
    a_env = DroneChaseVecEnv(name, n_envs=n, seed=6, sparse_lidar=True, host_threads=3)
    b_env = DroneChaseVecEnv(name, n_envs=n, seed=6, sparse_lidar=False)
    c_env = DroneChaseVecEnv(name, n_envs=n, seed=6, sparse_lidar=True, mapped_lidar=False, host_threads=2)
    assert a_env.pairs == (not level5) and not a_env.mapped and not c_env.mapped and not c_env.pairs and not b_env.mapped
    d_env = None if level5 else DroneChaseVecEnv(name, n_envs=n, seed=6, sparse_lidar=True, mapped_lidar=True)
    assert d_env is None or (d_env.mapped and not d_env.pairs)
    if not level5:
        a_env._pairs_fast = 64                 # most steps need the second copy of the change list: both paths run
    if d_env is not None:
        od = d_env.reset()
    assert a_env.d2h_bytes_per_step < b_env.d2h_bytes_per_step / 5
    oa, ob, oc = a_env.reset(), b_env.reset(), c_env.reset()
    rng = np.random.RandomState(2)
    marked = 0
    held = []
    for t in range(150):
        a = np.concatenate([rng.uniform(-1, 1, (n, 3)), rng.uniform(0, 1, (n, 1))], axis=1).astype(np.float32)
        oa, ra, da, ia = a_env.step(a)
        ob, rb, db, ib = b_env.step(a)
        oc, rc, dc_, _ = c_env.step(a)
        for k in ob:
            assert np.array_equal(oa[k], ob[k]), f"step {t}: {k} (default transfer)"
            assert np.array_equal(oc[k], ob[k]), f"step {t}: {k} (host scatter)"
        if d_env is not None:
            od, rd, dd, _ = d_env.step(a)
            for k in ob:
                assert np.array_equal(od[k], ob[k]), f"step {t}: {k} (mapped mirror)"
            assert np.array_equal(rd, rb) and np.array_equal(dd, db)
        assert np.array_equal(ra, rb) and np.array_equal(da, db) and np.array_equal(rc, rb) and np.array_equal(dc_, db)
        for i in np.nonzero(da)[0]:                # lazily built terminal observations: same rows on both paths
            ta, tb = ia[int(i)]["terminal_observation"], ib[int(i)]["terminal_observation"]
            assert set(ta) == set(tb)
            for k in tb:
                assert np.array_equal(ta[k], tb[k]), f"step {t} env {i}: terminal {k}"
        lk = "stacked_spheres" if level5 else "lidar"
        marked += int((oa[lk] < 1).sum())
        held.append((oa[lk], oa[lk].copy()))
        if len(held) > 1:                      # the arrays of step t-1 are still intact while step t is handed out
            arr, snap = held.pop(0)
            assert np.array_equal(arr, snap)
    assert marked > 1000
    a_env.close(); b_env.close(); c_env.close()
    if d_env is not None:
        d_env.close()


def test_shipped_pipeline_returns_a_monitored_gpu_vec_env(monkeypatch):
    """compat/core/rl_framework/utils/pipeline.py (reference pipeline.py:32-61): the call every training app makes returns
    VecMonitor(DroneChaseVecEnv) -- here with the fallback VecMonitor, SB3 being absent -- and the monitor's episode
    records agree with the simulator's own episode statistics."""
    import os
    import sys
    compat = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "compat")
    monkeypatch.syspath_prepend(compat)
    monkeypatch.setenv("DRONECHASE_B200_ENVS", "512")
    for m in [k for k in sys.modules if k.split(".")[0] in ("core", "threatengage", "threatsense")]:
        monkeypatch.delitem(sys.modules, m)
    from threatengage.rl_framework.utils.pipeline import ReinforcementLearningPipeline            # the apps' import path
    from threatengage.environments.level4.exp02_vFinal_environment import Exp02vFinalEnvironment as level4
    venv = ReinforcementLearningPipeline.create_vectorized_environment(environment=level4, env_kwargs={"rl_frequency": 15, "learning_rate": 1e-4})
    assert type(venv).__name__ == "VecMonitor" and venv.num_envs == 512 and type(venv.venv).__name__ == "DroneChaseVecEnv"
    obs = venv.reset()
    assert obs["lidar"].shape == (512, 3, 13, 26)
    rng = np.random.RandomState(0)
    n_ep, ret, length = 0, 0.0, 0
    for t in range(200):
        a = np.concatenate([rng.uniform(-1, 1, (512, 3)), rng.uniform(0, 1, (512, 1))], axis=1).astype(np.float32)
        obs, rew, dones, infos = venv.step(a)
        for i in np.nonzero(dones)[0]:
            ep = infos[int(i)]["episode"]
            assert "terminal_observation" in infos[int(i)]
            n_ep += 1; ret += ep["r"]; length += ep["l"]
    stats = venv.venv.sim.stats.cpu().numpy()
    assert n_ep >= 5 and n_ep == int(stats[0]) and length == int(stats[2])
    assert abs(ret - stats[1]) <= 1e-3 * max(1.0, abs(stats[1]))
    venv.close()

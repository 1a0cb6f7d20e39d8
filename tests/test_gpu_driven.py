"""Level4 tasks whose wingmen are flown by policies inside the task, on CUDA through the C ABI (needs a B200):
``EvaluationEnvironment`` + ``Evaluation_Task`` and ``Exp05vFinalEnvironment`` + ``Exp05_vFinal_Task``.

  * the recordings of the reference's OWN classes (tests/golden/l4eval_*.npz, l4exp05_*.npz) replayed through the f64
    build: the observation every policy-driven wingman is handed at on_step_start (dc_lw_observe: sphere cells exact,
    floats 1e-6), the env observation, terminated, the per-wingman info rows;
  * f64 closed loop over a batch against oracle/eval_oracle.py with auto-reset (exact events, kills, waves);
  * f32 (product) closed loop under the margin protocol of test_gpu_stage03.
The policies are oracle/eval_policy.pilot on both sides (a fixed function of the whole observation dict), evaluated on the
host from the tensors dc_lw_observe produced; the shared last_action chain is dronechase_b200.drivers.TaskDrivers'."""
import dataclasses
import glob
import os

import numpy as np
import pytest
import torch

from oracle.eval_oracle import EXP05, DrivenOracle, evaluation_config
from oracle.eval_policy import pilot
from oracle.make_golden_eval import case_config
from tests.util import kite_actions, load_recording

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
_DRV = {"agent": "legacy", "nn": "nn", "bt": "bt", "stop": "stop", "nn_ally": "nn_ally"}


def _task_config(cfg, **kw):
    from dronechase_b200 import TaskConfig
    return TaskConfig(n_lw=cfg.n_lw, n_lm=cfg.n_lm, munition=cfg.munition, born_radius=cfg.born_radius, initial_round=cfg.initial_round,
                      step_increment=cfg.step_increment, max_step=cfg.max_step, lw_driver=tuple(_DRV[d] for d in cfg.drivers),
                      eval_task=cfg.task == "evaluation", time_is_limited=cfg.time_limited, noise_ratio=cfg.noise_ratio, **kw)


def _pilot_policy(salt):
    def call(obs):
        a = pilot(obs["lidar"].cpu().numpy(), obs["inertial_data"].cpu().numpy(), obs["last_action"].cpu().numpy(), salt)
        return torch.from_numpy(a).to(obs["lidar"].device)
    return call


def _make(cfg, E, seed, precision, auto_reset, env_offset=0, salts=None):
    from dronechase_b200 import BatchedThreatEngageEnv
    from dronechase_b200.drivers import TaskDrivers
    salts = list(salts) if salts is not None else [0.0] * cfg.n_lw
    env = BatchedThreatEngageEnv(_task_config(cfg), n_envs=E, seed=seed, device=0, env_offset=env_offset, auto_reset=auto_reset,
                                 precision=precision, with_ids=True)
    drv = TaskDrivers(env, {j: _pilot_policy(salts[j]) for j in env.cfg.policy_slots})
    orc = DrivenOracle(cfg, E, seed=seed, env_offset=env_offset, auto_reset=auto_reset, salts=salts)
    return env, drv, orc


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "l4e*.npz"))), ids=lambda p: os.path.basename(p)[:-4])
def test_driven_golden_replay_through_cuda(path):
    rec = load_recording(path)
    seed, env_index, n_steps, _ = (int(v) for v in rec["meta"])
    kind, spec, extra = str(rec["kind"]), tuple(str(s) for s in rec["drivers"]), eval(str(rec["extra"]))
    cfg, _ = case_config(kind, spec, extra)
    cfg = dataclasses.replace(cfg, noise_ratio=float(rec["noise_ratio"]))
    from dronechase_b200 import BatchedThreatEngageEnv
    from dronechase_b200.drivers import TaskDrivers
    env = BatchedThreatEngageEnv(_task_config(cfg), n_envs=1, seed=seed, device=0, env_offset=env_index, auto_reset=False,
                                 precision="f64", with_ids=True)
    drv = TaskDrivers(env, {j: _pilot_policy(float(rec["salts"][j])) for j in env.cfg.policy_slots})
    obs = env.reset(); drv.reset()
    k = 0

    def check_obs(tag):
        got = obs["lidar"].cpu().numpy()[0]
        assert np.array_equal(got < 1, rec["lidar"][k] < 1), f"{tag}: marked cells of the env observation"
        assert np.abs(got - rec["lidar"][k]).max() < 1e-6, f"{tag}: sphere"
        assert np.abs(obs["inertial_data"].cpu().numpy()[0] - rec["inertial"][k]).max() < 1e-6, f"{tag}: inertial"
        assert np.abs(obs["last_action"].cpu().numpy()[0] - rec["last_action"][k]).max() < 1e-6, f"{tag}: last_action"
    check_obs("reset"); k += 1
    n_calls = 0
    for t in range(n_steps):
        tag = f"{os.path.basename(path)} step {t}"
        last_before = drv.last_action.clone()
        drv.serve()
        lw = env.lw_obs
        present = lw["present"].cpu().numpy()[0]
        want = rec["nn_called"][t]
        assert np.array_equal(present[list(env.cfg.policy_slots)], want[list(env.cfg.policy_slots)]), f"{tag}: wingmen served"
        chain = last_before[0].cpu().numpy()
        for j in env.cfg.policy_slots:
            if not want[j]:
                continue
            got = lw["lidar"][0, j].cpu().numpy()
            assert np.array_equal(got < 1, rec["nn_lidar"][t, j] < 1), f"{tag}: wingman {j} policy sphere cells"
            assert np.abs(got - rec["nn_lidar"][t, j]).max() < 1e-6, f"{tag}: wingman {j} policy sphere"
            assert np.abs(lw["inertial_data"][0, j].cpu().numpy() - rec["nn_inertial"][t, j]).max() < 1e-6, f"{tag}: wingman {j} inertial"
            assert np.abs(chain - rec["nn_last_action"][t, j]).max() < 1e-6, f"{tag}: wingman {j} shared last_action"
            assert np.abs(env.lw_actions[0, j].cpu().numpy() - rec["nn_action"][t, j]).max() < 1e-6, f"{tag}: wingman {j} action"
            chain = env.lw_actions[0, j].cpu().numpy()
            n_calls += 1
        obs, rew, done, info = env.step(torch.from_numpy(rec["actions"][t][None].astype(np.float32)).cuda())
        assert abs(float(rew[0]) - rec["reward"][t]) <= 1e-3 + 1e-6 * abs(rec["reward"][t]), f"{tag}: reward"
        assert bool(done[0]) == bool(rec["done"][t]), f"{tag}: terminated"
        inf = env.info.cpu().numpy()[0]
        if kind == "evaluation":
            li = env.lw_info.cpu().numpy()[0]
            alive = rec["lw_alive"][t]
            assert np.array_equal(li[:, 1].astype(bool), alive), f"{tag}: armed wingmen"
            assert np.array_equal(li[alive, 0], rec["lw_kills"][t][alive]) and np.array_equal(li[alive, 2], rec["lw_munitions"][t][alive]), f"{tag}: lw rows"
            if alive.any():
                assert int(inf[3]) == int(rec["wave"][t]) and int(inf[5]) == int(rec["step"][t]), f"{tag}: wave / step"
        else:
            assert [int(v) for v in inf[:4]] == [int(v) for v in rec["info4"][t]], f"{tag}: info"
        check_obs(tag); k += 1
        if done[0]:
            obs = env.reset(); drv.reset()
            check_obs(tag + " reset"); k += 1
    assert n_calls > 100
    env.close()


CLOSED = [("exp05", EXP05), ("eval_nn_bt", evaluation_config(("nn", "bt"), time_limited=True, max_step=150, step_increment=40)),
          ("eval_2nn", evaluation_config(("nn", "nn"), initial_round=2, time_limited=True, max_step=110, step_increment=30)),
          ("eval_bt_nn_stop", evaluation_config(("bt", "nn", "stop"), munition=8, time_limited=True, max_step=130, step_increment=30))]


@pytest.mark.parametrize("name,cfg", CLOSED, ids=[c[0] for c in CLOSED])
def test_driven_closed_loop_f64_exact(name, cfg):
    E, K = 24, 260
    salts = [0.0, 0.52, 0.03][:cfg.n_lw]                 # the second policy rams: wingmen die, episodes end
    env, drv, orc = _make(cfg, E, seed=23, precision="f64", auto_reset=True, salts=salts)
    obs = env.reset(); drv.reset(); ref = orc.reset()
    rng = np.random.RandomState(3)
    kills = episodes = served = 0
    for t in range(K):
        a = kite_actions(orc, rng, ram=(t > 120)) if cfg.task != "evaluation" else np.zeros((E, 4), dtype=np.float32)
        drv.serve()
        obs, rew, done, info = env.step(torch.from_numpy(a).cuda())
        if env._c.auto_reset:
            drv.last_action = torch.where(env.done.bool()[:, None], torch.zeros_like(drv.last_action), drv.last_action)
        ref, r_ref, d_ref, i_ref = orc.step(a.astype(np.float64))
        tag = f"{name} step {t}"
        called = orc.nn_obs["called"]
        assert np.array_equal(env.lw_obs["present"].cpu().numpy()[:, list(env.cfg.policy_slots)], called[:, list(env.cfg.policy_slots)]), f"{tag}: served"
        for j in env.cfg.policy_slots:
            m = called[:, j]
            got = env.lw_obs["lidar"][:, j].cpu().numpy()[m]
            assert np.array_equal(got < 1, orc.nn_obs["lidar"][m, j] < 1), f"{tag}: wingman {j} sphere cells"
            assert np.abs(got - orc.nn_obs["lidar"][m, j]).max(initial=0.0) < 1e-6, f"{tag}: wingman {j} sphere"
            assert np.abs(env.lw_obs["inertial_data"][:, j].cpu().numpy()[m] - orc.nn_obs["inertial"][m, j]).max(initial=0.0) < 1e-6, f"{tag}: wingman {j} inertial"
            assert np.abs(env.lw_actions[:, j].cpu().numpy()[m] - orc.nn_obs["action"][m, j]).max(initial=0.0) < 1e-6, f"{tag}: wingman {j} action"
            served += int(m.sum())
        assert np.array_equal(done.cpu().numpy().astype(bool), d_ref), f"{tag}: terminated"
        assert np.allclose(rew.cpu().numpy(), r_ref, rtol=1e-6, atol=1e-5), f"{tag}: reward"
        inf = env.info.cpu().numpy()
        if cfg.task == "evaluation":
            li = env.lw_info.cpu().numpy()
            assert np.array_equal(li[..., 1].astype(bool), i_ref["lw_alive"]) and np.array_equal(li[..., 0], i_ref["lw_kills"]), f"{tag}: lw rows"
            assert np.array_equal(li[..., 2], i_ref["lw_munitions"]) and np.array_equal(inf[:, 3], i_ref["current_wave"]), f"{tag}: munitions / wave"
            kills = max(kills, int(i_ref["lw_kills"].sum(axis=1).max()))
        else:
            for col, key in ((0, "agent_kills"), (1, "allies_kills"), (2, "deads"), (3, "current_wave")):
                assert np.array_equal(inf[:, col], i_ref[key]), f"{tag}: {key}"
            kills = max(kills, int((i_ref["agent_kills"] + i_ref["allies_kills"]).max()))
        assert np.array_equal(obs["lidar"].cpu().numpy() < 1, ref["lidar"] < 1), f"{tag}: env observation cells"
        assert np.allclose(obs["lidar"].cpu().numpy(), ref["lidar"], atol=1e-6) and np.allclose(obs["inertial_data"].cpu().numpy(), ref["inertial_data"], atol=1e-6), tag
        assert np.allclose(obs["last_action"].cpu().numpy(), ref["last_action"], atol=1e-6), f"{tag}: last_action"
        episodes += int(d_ref.sum())
    st = env.get_state()
    # (a policy's float32 action goes through a float32 norm / division in convert_command_to_setpoint: numpy's and the
    # kernel's can differ in the last float32 bit of a setpoint, 1e-8 m/s -- positions agree to 1e-6 m, the stated tolerance)
    assert np.array_equal(st["armed"], orc.armed) and np.abs(st["pos"] - orc.pos)[orc.armed].max() < 1e-6
    assert np.array_equal(st["spawn_ctr"], orc.spawn_ctr) and np.array_equal(st["hit_ctr"], orc.hit_ctr)
    assert kills >= 1 and episodes >= 1 and served > 0.5 * E * K, (kills, episodes, served)
    env.close()


def test_driven_closed_loop_f32():
    cfg = evaluation_config(("nn", "bt"), time_limited=True, max_step=200, step_increment=50)
    E, K, MARGIN = 64, 160, 2e-4
    env, drv, orc = _make(cfg, E, seed=29, precision="f32", auto_reset=True)
    env.reset(); drv.reset(); orc.reset()
    excused = np.zeros(E, dtype=bool)
    cells_cmp = cells_bad = 0
    for t in range(K):
        drv.serve()
        obs, rew, done, info = env.step(None)
        drv.last_action = torch.where(env.done.bool()[:, None], torch.zeros_like(drv.last_action), drv.last_action)
        orc.min_margin[:] = np.inf
        ref, r_ref, d_ref, i_ref = orc.step()
        excused |= orc.min_margin < MARGIN
        ok = ~excused
        assert np.array_equal(done.cpu().numpy().astype(bool)[ok], d_ref[ok]), f"step {t}: terminated"
        li = env.lw_info.cpu().numpy()
        assert np.array_equal(li[ok, :, 0], i_ref["lw_kills"][ok]) and np.array_equal(li[ok, :, 1].astype(bool), i_ref["lw_alive"][ok]), f"step {t}: lw rows"
        m = ok & orc.nn_obs["called"][:, 0]
        got, want = env.lw_obs["lidar"][:, 0].cpu().numpy()[m], orc.nn_obs["lidar"][m, 0]
        cells_cmp += int((want < 1).sum()); cells_bad += int(((got < 1) != (want < 1)).sum())
        assert np.abs(env.lw_obs["inertial_data"][:, 0].cpu().numpy()[m] - orc.nn_obs["inertial"][m, 0]).max(initial=0.0) < 5e-4, f"step {t}: inertial"
    assert excused.mean() < 0.1 and cells_cmp > 1000 and cells_bad < 0.02 * cells_cmp, (excused.mean(), cells_cmp, cells_bad)
    env.close()


class _NumpyModel:
    """Stands in for an SB3 model: predict(observation dict of numpy arrays, deterministic=True) -> (actions, None)."""
    def __init__(self, salt=0.0): self.salt = salt
    def predict(self, observation, deterministic=True):
        return pilot(observation["lidar"], observation["inertial_data"], observation["last_action"], self.salt), None


def test_facades_evaluation_and_exp05():
    from dronechase_b200.gym_env import EvaluationEnvironment, Exp05vFinalEnvironment
    configuration = {"drivers": [{"type": "nn", "name": "nn_1", "policy": _NumpyModel()}, {"type": "bt", "name": "bt_1"}],
                     "TIME_IS_LIMITED": True, "MAX_STEP": 60, "STEP_INCREMENT": 0}
    env = EvaluationEnvironment(configuration, GUI=False, rl_frequency=15, seed=3)
    obs, info = env.reset(0)
    assert info == {} and obs["lidar"].shape == (3, 13, 26) and not obs["last_action"].any()
    ended = False
    for t in range(70):
        obs, reward, terminated, truncated, info = env.step(np.zeros(1))
        assert reward == 0.0 and truncated is False and not obs["last_action"].any()
        assert set(info) <= {"nn_1", "bt_1"} and all(set(v) == {"lw_kills", "lw_alive", "lw_munitions", "current_wave", "step"} for v in info.values())
        assert all(v["step"] == t + 1 and v["lw_alive"] is True for v in info.values())
        if terminated:
            ended = True
            break
    assert ended and t >= 60                                  # the time limit ended it (step > MAX_STEP)
    env.close()
    env = Exp05vFinalEnvironment(dome_radius=20, rl_frequency=15, GUI=False, seed=4)
    with pytest.raises(RuntimeError):
        env.step(np.array([0.1, 0.0, 0.0, 0.5]))
    env.update_model(_NumpyModel(0.02))
    obs, info = env.reset()
    for t in range(25):
        obs, reward, terminated, truncated, info = env.step(np.array([0.3, -0.2, 0.1, 0.8], dtype=np.float32))
    assert set(info) == {"agent_kills", "allies_kills", "deads", "current_wave"} and np.allclose(obs["last_action"], [0.3, -0.2, 0.1, 0.8])
    # the ally really flies under its policy: its commanded action is the pilot's
    a = env.sim.lw_actions[0, 1].cpu().numpy()
    assert 0.29 <= a[3] <= 0.95 and np.abs(a[:3]).max() > 0.05
    env.close()


def test_evaluate_level4_matches_the_oracle_episode_table():
    """evaluate_level4 (the evaluation_exp0*_app_ready.py loop over one batch) on the f32 product build against the same
    bookkeeping done on the float64 oracle: kills / munitions / wave / last step per wingman and episode."""
    from dronechase_b200.evaluation import evaluate_level4
    configuration = {"drivers": [{"type": "nn", "name": "nn_1"}, {"type": "bt", "name": "bt_1"}], "TIME_IS_LIMITED": True,
                     "MAX_STEP": 90, "STEP_INCREMENT": 20}
    E, N = 8, 16
    rows, cols = evaluate_level4(configuration, n_episodes=N, n_envs=E, seed=31, policies={0: _NumpyModel()})
    assert cols == ["kills", "alive", "munitions", "wave", "step", "name", "episode"]
    cfg = evaluation_config(("nn", "bt"), time_limited=True, max_step=90, step_increment=20)
    orc = DrivenOracle(cfg, E, seed=31, auto_reset=True)
    orc.reset()
    quota, counts = N // E, np.zeros(E, dtype=np.int64)
    last = np.zeros((E, 2, 5), dtype=np.int64); seen = np.zeros((E, 2), dtype=bool)
    eps = []
    while counts.min() < quota:
        _, _, d, info = orc.step()
        armed = info["lw_alive"]
        row = np.stack([info["lw_kills"], armed.astype(np.int64), info["lw_munitions"], np.repeat(info["current_wave"][:, None], 2, 1),
                        np.repeat(info["step"][:, None], 2, 1)], axis=-1)
        last = np.where(armed[..., None], row, last); seen |= armed
        for e in np.nonzero(d)[0]:
            if counts[e] < quota:
                eps.append((int(counts[e]), int(e), [[int(v) for v in last[e, j]] + [("nn_1", "bt_1")[j]] for j in range(2) if seen[e, j]]))
                counts[e] += 1
            last[e] = 0; seen[e] = False
    eps.sort(key=lambda x: (x[0], x[1]))
    want = [[r[0], bool(r[1]), r[2], r[3], r[4], r[5], i + 1] for i, (_, _, rr) in enumerate(eps[:N]) for r in rr]
    assert len(rows) == len(want) and len({r[6] for r in rows}) == N
    # float32 vs float64 trajectories: the discrete columns agree on (almost) every row
    same = sum(1 for a, b in zip(rows, want) if a == b)
    assert same >= 0.8 * len(want), (same, len(want))

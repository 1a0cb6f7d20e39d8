"""threatsense Level5DumbMultiObs (the data-collection env: every armed wingman observes, all fly the behaviour tree)
on CUDA through the C ABI -- needs a B200.

  * recordings of the reference's OWN Level5DumbMultiObs + Level5DumbMultiObjectTask (tests/golden/l5dumb_*.npz)
    replayed through the f64 build: observing wingmen, validity masks, marked cells, kills / waves / terminations exact,
    stacks, inertial vectors and teacher actions to 1e-6;
  * f64 closed loop against oracle/level5_oracle.py (LEVEL5_DUMB) over a batch with auto-reset: exact.
"""
import dataclasses
import glob
import os

import numpy as np
import pytest
import torch

from oracle.level5_oracle import LEVEL5_DUMB, Level5Oracle
from tests.util import load_recording

pytestmark = pytest.mark.gpu


def _cmp_multi(mo, ref, tag, atol=1e-6):
    pres = mo["present"].cpu().numpy()
    assert np.array_equal(pres, ref["present"]), f"{tag}: observing wingmen"
    assert np.array_equal(mo["validity_mask"].cpu().numpy()[pres], ref["validity_mask"][pres]), f"{tag}: validity masks"
    got, want = mo["stacked_spheres"].cpu().numpy()[pres], ref["stacked_spheres"][pres]
    assert np.array_equal(got < 1, want < 1), f"{tag}: stacks mark different cells"
    assert np.abs(got - want).max(initial=0.0) <= atol, f"{tag}: stacks"
    assert np.abs(mo["inertial_data"].cpu().numpy()[pres] - ref["inertial_data"][pres]).max(initial=0.0) <= atol, f"{tag}: inertial"
    assert np.abs(mo["last_action"].cpu().numpy()[pres] - ref["last_action"][pres]).max(initial=0.0) <= atol, f"{tag}: teacher actions"


def test_dumb_multiobs_golden_replay_through_cuda(golden_dir):
    from dronechase_b200 import BatchedThreatEngageEnv, preset
    paths = sorted(glob.glob(os.path.join(golden_dir, "l5dumb_*.npz")))
    assert len(paths) >= 2
    for path in paths:
        rec = load_recording(path)
        seed, env_index, n_steps, _ = (int(v) for v in rec["meta"])
        env = BatchedThreatEngageEnv(preset("level5_dumb_multiobs", noise_ratio=float(rec["noise_ratio"]),
                                            step_increment=int(rec["step_increment"])), n_envs=1, seed=seed,
                                     env_offset=env_index, auto_reset=False, precision="f64")
        env.reset()
        k = 0

        def check(tag):
            mo = env.multi_obs
            pres = rec["present"][k]
            assert np.array_equal(mo["present"].cpu().numpy()[0], pres), f"{tag}: observing wingmen"
            assert np.array_equal(mo["validity_mask"].cpu().numpy()[0][pres], rec["mask"][k][pres]), f"{tag}: validity masks"
            got, want = mo["stacked_spheres"].cpu().numpy()[0][pres], rec["stacked"][k][pres]
            assert np.array_equal(got < 1, want < 1), f"{tag}: marked cells"
            assert np.abs(got - want).max(initial=0.0) < 1e-6, f"{tag}: stacks"
            assert np.abs(mo["inertial_data"].cpu().numpy()[0][pres] - rec["inertial"][k][pres]).max(initial=0.0) < 1e-6, f"{tag}: inertial"
            assert np.abs(mo["last_action"].cpu().numpy()[0][pres] - rec["teacher_actions"][k][pres]).max(initial=0.0) < 1e-6, f"{tag}: teacher actions"
        check(f"{path} reset"); k += 1
        for t in range(n_steps):
            _, rew, done, info = env.step(None)
            assert abs(float(rew[0]) - rec["reward"][t]) <= 1e-3 + 1e-6 * abs(rec["reward"][t]), f"{path} step {t}: reward"
            assert bool(done[0]) == bool(rec["done"][t]), f"{path} step {t}: done"
            inf = env.info.cpu().numpy()[0]
            assert [int(v) for v in inf[:4]] == [int(v) for v in rec["info"][t]], f"{path} step {t}: info {inf[:4]} vs {rec['info'][t]}"
            check(f"{path} step {t}"); k += 1
            if done[0]:
                env.reset()
                check(f"{path} reset after step {t}"); k += 1
        env.close()


def test_dumb_multiobs_closed_loop_f64_exact():
    from dronechase_b200 import BatchedThreatEngageEnv, preset
    E, K = 6, 150
    kw = {"step_increment": 3, "max_step": 60}                  # short episodes: time-outs + auto-resets inside K steps
    env = BatchedThreatEngageEnv(preset("level5_dumb_multiobs", **kw), n_envs=E, seed=41, device=0, auto_reset=True,
                                 precision="f64", sub_batches=2)
    orc = Level5Oracle(dataclasses.replace(LEVEL5_DUMB, **kw), E, seed=41, auto_reset=True)
    env.reset(); ref = orc.reset()
    _cmp_multi(env.multi_obs, ref, "reset")
    resets = kills = 0
    for t in range(K):
        _, rew, done, info = env.step(None)
        ref, r_ref, d_ref, i_ref = orc.step(np.zeros((E, 4)))
        assert np.array_equal(done.cpu().numpy().astype(bool), d_ref), f"step {t}: terminated flags"
        inf = env.info.cpu().numpy()
        for col, key in ((0, "agent_kills"), (1, "allies_kills"), (2, "deads"), (3, "current_wave")):
            assert np.array_equal(inf[:, col], i_ref[key]), f"step {t}: {key}"
        assert np.allclose(rew.cpu().numpy(), r_ref, rtol=1e-6, atol=1e-5), f"step {t}: reward"
        _cmp_multi(env.multi_obs, ref, f"step {t}")
        resets += int(d_ref.sum()); kills = max(kills, int((i_ref["agent_kills"] + i_ref["allies_kills"]).max()))
    assert resets >= 2 and kills >= 1, f"scenario too tame: {resets} episodes, {kills} kills"
    env.close()


def test_dumb_multiobs_collection(tmp_path):
    """collect_data_multiobs: the flow of apps/threatsense_runner/collect_and_save.py (drop rows without a valid sphere,
    student-only parts of N samples) over the batched env; f32 product build."""
    from dronechase_b200 import BatchedThreatEngageEnv
    from dronechase_b200.io_data import DatasetWriter, MultiFileDataset, _open_part, collect_data_multiobs
    env = BatchedThreatEngageEnv("level5_dumb_multiobs", n_envs=64, seed=2, auto_reset=True)
    with DatasetWriter(str(tmp_path), samples_per_file=1000, backend="npz", file_stem="part") as w:
        res = collect_data_multiobs(env, w, max_observations_collected=2500)
    assert res["observations"] == 2500
    ds = MultiFileDataset(str(tmp_path))
    assert len(ds) == 2500 and len(ds.file_paths) == 3
    part = _open_part(ds.file_paths[0])
    assert set(part) == {"student/stacked_spheres", "student/validity_mask", "student/inertial_data", "student/last_action",
                         "teacher_actions"}
    assert part["student/validity_mask"].any(axis=1).all()
    assert np.array_equal(part["student/last_action"], part["teacher_actions"])       # level5_dumb_multiobs.py:141-146
    n = np.linalg.norm(part["teacher_actions"][:, :3], axis=1)
    assert (np.abs(n - 1) < 1e-5).mean() > 0.9                        # behaviour-tree commands: unit direction + speed
    env.close()


def test_dumb_multiobs_closed_loop_f32():
    """The f32 product build against the float64 oracle: draws, masks and observing wingmen are integer logic and stay
    exact; an env whose oracle reports a predicate within MARGIN of its threshold is excused from then on; a float32
    pose may move a hit across a cell border (< 1 % of the marked cells)."""
    from dronechase_b200 import BatchedThreatEngageEnv, preset
    E, K, MARGIN = 12, 120, 2e-4
    env = BatchedThreatEngageEnv(preset("level5_dumb_multiobs"), n_envs=E, seed=9, device=0, auto_reset=True, precision="f32")
    orc = Level5Oracle(LEVEL5_DUMB, E, seed=9, auto_reset=True)
    env.reset(); orc.reset()
    excused = np.zeros(E, dtype=bool)
    cells_cmp = cells_bad = 0
    for t in range(K):
        _, rew, done, info = env.step(None)
        orc.min_margin[:] = np.inf
        ref, r_ref, d_ref, i_ref = orc.step(np.zeros((E, 4)))
        excused |= orc.min_margin < MARGIN
        ok = ~excused
        inf = env.info.cpu().numpy()
        assert np.array_equal(done.cpu().numpy().astype(bool)[ok], d_ref[ok]), f"step {t}: terminated flags"
        for col, key in ((0, "agent_kills"), (1, "allies_kills"), (2, "deads"), (3, "current_wave")):
            assert np.array_equal(inf[ok, col], i_ref[key][ok]), f"step {t}: {key}"
        mo = {k: v.cpu().numpy() for k, v in env.multi_obs.items()}
        assert np.array_equal(mo["present"][ok], ref["present"][ok]), f"step {t}: observing wingmen"
        sel = ref["present"] & ok[:, None]
        assert np.array_equal(mo["validity_mask"][sel], ref["validity_mask"][sel]), f"step {t}: validity masks"
        assert np.abs(mo["inertial_data"][sel] - ref["inertial_data"][sel]).max(initial=0.0) < 5e-4, f"step {t}: inertial"
        assert np.abs(mo["last_action"][sel] - ref["last_action"][sel]).max(initial=0.0) < 2e-3, f"step {t}: teacher actions"
        got, want = mo["stacked_spheres"][sel], ref["stacked_spheres"][sel]
        cells_cmp += int((want < 1).sum()); cells_bad += int(((got < 1) != (want < 1)).sum())
        both = (got < 1) & (want < 1)
        assert np.abs(got - want)[both].max(initial=0.0) < 5e-4, f"step {t}: stacked distances"
    assert excused.mean() < 0.2, f"too many envs excused: {excused.mean()}"
    assert cells_cmp > 1000 and cells_bad < 0.01 * cells_cmp, f"{cells_bad} of {cells_cmp} marked cells differ"
    env.close()


def test_eval2bt_golden_replay_and_closed_loop(golden_dir):
    """Level52BTEvaluationEnvironment (two behaviour-tree wingmen, no observation, no reward): the reference's own
    recordings replayed through the f64 build, then an f64 closed loop over a batch against the oracle."""
    from dronechase_b200 import BatchedThreatEngageEnv, preset
    from dronechase_b200.gym_env import Level52BTEvaluationEnvironment
    from oracle.level5_oracle import LEVEL5_EVAL2BT
    from tests.util import load_recording
    paths = sorted(glob.glob(os.path.join(golden_dir, "l5eval2bt_*.npz")))
    assert len(paths) >= 2
    for path in paths:
        rec = load_recording(path)
        seed, env_index, n_steps, _ = (int(v) for v in rec["meta"])
        env = BatchedThreatEngageEnv(preset("level5_eval_2bt", noise_ratio=float(rec["noise_ratio"]), max_step=int(rec["max_step"])),
                                     n_envs=1, seed=seed, env_offset=env_index, auto_reset=False, precision="f64")
        env.reset()
        k = 0

        def check(tag):
            st = env.get_state()
            assert np.array_equal(st["armed"][0], rec["armed"][k]), f"{tag}: armed flags"
            a = rec["armed"][k]
            assert np.abs(st["pos"][0][a] - rec["pos"][k][a]).max(initial=0.0) < 1e-6, f"{tag}: positions"
        check(f"{path} reset"); k += 1
        for t in range(n_steps):
            _, rew, done, info = env.step(None)
            assert float(rew[0]) == 0.0 and bool(done[0]) == bool(rec["done"][t]), f"{path} step {t}: reward / done"
            inf = env.info.cpu().numpy()[0]
            want = [int(rec["kills"][t][0]), int(rec["kills"][t][1]), int(rec["info"][t][0]), int(rec["info"][t][1])]
            assert [int(v) for v in inf[:4]] == want, f"{path} step {t}: info {inf[:4]} vs {want}"
            check(f"{path} step {t}"); k += 1
            if done[0]:
                env.reset()
                check(f"{path} reset after step {t}"); k += 1
        env.close()
    E, K = 24, 120
    kw = {"max_step": 50}
    env = BatchedThreatEngageEnv(preset("level5_eval_2bt", **kw), n_envs=E, seed=77, device=0, auto_reset=True, precision="f64")
    orc = Level5Oracle(dataclasses.replace(LEVEL5_EVAL2BT, **kw), E, seed=77, auto_reset=True)
    env.reset(); orc.reset()
    resets = 0
    for t in range(K):
        _, rew, done, info = env.step(None)
        _, r_ref, d_ref, i_ref = orc.step(np.zeros((E, 4)))
        assert np.array_equal(done.cpu().numpy().astype(bool), d_ref) and not rew.cpu().numpy().any(), f"step {t}"
        inf = env.info.cpu().numpy()
        for col, key in ((0, "agent_kills"), (1, "allies_kills"), (2, "deads"), (3, "current_wave")):
            assert np.array_equal(inf[:, col], i_ref[key]), f"step {t}: {key}"
        resets += int(d_ref.sum())
    st = env.get_state()
    assert np.array_equal(st["armed"], orc.armed) and np.abs(st["pos"] - orc.pos)[orc.armed].max() < 1e-7
    assert resets >= E
    env.close()
    e1 = Level52BTEvaluationEnvironment(GUI=False, seed=1)
    obs, info = e1.reset()
    assert obs == {} and set(info) == {"kills_per_drone", "deads", "current_wave"}
    obs, r, term, trunc, info = e1.step(np.zeros(4))
    assert obs == {} and r == 0.0 and trunc is False and info["kills_per_drone"][1]["type"] == "BT"
    e1.close()


def test_evaluate_2bt_matches_the_oracle_episode_table():
    """evaluate_2bt (apps/threatsense_runner/evaluation_2bt.py over the batch): the per-episode kill table of the f32
    product build equals the one the float64 oracle produces for the same seed (short episodes: MAX_STEP 40)."""
    from dronechase_b200.evaluation import episodes_from_steps, evaluate_2bt, select_rows, summarise
    from oracle.level5_oracle import LEVEL5_EVAL2BT
    E, N, kw = 16, 40, {"max_step": 40}
    rows, raw, stats = evaluate_2bt(n_episodes=N, n_envs=E, seed=5, **kw)
    assert len(rows) == N and set(raw) == {"loyalwingman_0", "loyalwingman_1", "total_kills"}
    orc = Level5Oracle(dataclasses.replace(LEVEL5_EVAL2BT, **kw), E, seed=5, auto_reset=True)
    orc.reset()
    want, t, quota, counts = [], 0, -(-N // E), np.zeros(E, dtype=np.int64)
    while counts.min() < quota:
        _, _, d_ref, i_ref = orc.step(np.zeros((E, 4)))
        t += 1
        info = np.zeros((E, 8), dtype=np.int64)
        info[:, 0], info[:, 1], info[:, 2], info[:, 3] = i_ref["agent_kills"], i_ref["allies_kills"], i_ref["deads"], i_ref["current_wave"]
        counts = episodes_from_steps(d_ref, info, t, want, quota, counts=counts)
    want = select_rows(want, N)
    # fixed quota per env (no length bias): every env contributes ceil(N / E) episodes before the cut
    assert max(r["episode"] for r in rows) == quota - 1
    key = lambda r: (r["episode"], r["env"], r["step"], r["loyalwingman_0"], r["loyalwingman_1"], r["deads"], r["current_wave"])   # noqa: E731
    assert [key(r) for r in rows] == [key(r) for r in want]
    assert stats["mean"]["total_kills"] == summarise(want)[1]["mean"]["total_kills"]

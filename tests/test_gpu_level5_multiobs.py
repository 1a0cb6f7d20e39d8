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

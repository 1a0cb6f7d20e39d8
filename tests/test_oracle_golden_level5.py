"""oracle/level5_oracle.py replayed against trajectories recorded from the reference's OWN
Level5C1FusionEnvironment / Level5C1FusionTask / FusedLIDAR / LiDARBufferManager code
(oracle/make_golden_level5.py).  Same tolerances as test_oracle_golden.py: float64 both sides -> 1e-9 on
reward/positions, exact flags / counters / hit ids / validity masks / drawn (publisher, age) pairs, 1e-6 on
the float32 observation tensors."""
import dataclasses
import glob
import os

import numpy as np
import pytest

from oracle.level5_oracle import LEVEL5_C1, LEVEL5_FUSION, Level5Oracle
from tests.util import load_recording

CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "level5_*.npz")))
# recordings of Level5FusionEnvironment (base Level5Environment + Level5FusionTask), oracle/make_golden_level5.py
FUSION_CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "l5fusion_*.npz")))


def _check_obs(rec, k, obs, orc, tag):
    assert (rec["armed"][k] == orc.armed[0]).all(), f"{tag}: armed flags differ"
    assert np.abs(rec["pos"][k] - orc.pos[0]).max() <= 1e-9, f"{tag}: positions differ"
    assert (rec["ammo"][k] == orc.ammo[0, :orc.cfg.n_lw]).all(), f"{tag}: ammunition differs"
    for name, key in (("inertial", "inertial_data"), ("last_action", "last_action"), ("sphere", "lidar")):
        d = np.abs(rec[name][k].astype(np.float64) - obs[key][0].astype(np.float64)).max()
        assert d <= 1e-6, f"{tag}: {name} differs by {d}"
    if not rec["was_reset"][k] and orc.armed[0, orc.agent[0]]:
        assert (rec["ids"][k] == orc.lidar_ids[0]).all(), f"{tag}: LiDAR hit ids differ"
        assert (rec["chosen"][k] == orc.chosen[0]).all(), f"{tag}: drawn (publisher, age) {orc.chosen[0].tolist()} vs {rec['chosen'][k].tolist()}"
    assert (rec["mask"][k] == obs["validity_mask"][0]).all(), f"{tag}: validity mask {obs['validity_mask'][0]} vs {rec['mask'][k]}"
    got, want = obs["stacked_spheres"][0], rec["stacked"][k]
    assert np.array_equal(got < 1, want < 1), f"{tag}: stacked spheres mark different cells"
    assert np.abs(got.astype(np.float64) - want.astype(np.float64)).max() <= 1e-6, f"{tag}: stacked spheres"
    if "stacked_student" in rec.files:       # info["student_observation"]: the second compute_observation call of the step
        if not rec["was_reset"][k] and orc.armed[0, orc.agent[0]]:
            assert (rec["chosen_student"][k] == orc.student_chosen[0]).all(), \
                f"{tag}: student draws {orc.student_chosen[0].tolist()} vs {rec['chosen_student'][k].tolist()}"
        assert (rec["mask_student"][k] == obs["student_validity_mask"][0]).all(), f"{tag}: student validity mask"
        got, want = obs["student_stacked_spheres"][0], rec["stacked_student"][k]
        assert np.array_equal(got < 1, want < 1), f"{tag}: student stack marks different cells"
        assert np.abs(got.astype(np.float64) - want.astype(np.float64)).max() <= 1e-6, f"{tag}: student stack"


@pytest.mark.parametrize("path", CASES + FUSION_CASES, ids=[os.path.basename(p)[:-4] for p in CASES + FUSION_CASES])
def test_level5_oracle_matches_reference_recording(path):
    rec = load_recording(path)
    seed, env_index, n_steps, _ = (int(v) for v in rec["meta"])
    base = LEVEL5_FUSION if os.path.basename(path).startswith("l5fusion_") else LEVEL5_C1
    cfg = dataclasses.replace(base, noise_ratio=float(rec["noise_ratio"]))
    orc = Level5Oracle(cfg, 1, seed=seed, env_offset=env_index)
    assert int(orc.agent[0]) == int(rec["agent_slot"])
    obs = orc.reset()
    k = 0
    _check_obs(rec, k, obs, orc, "reset"); k += 1
    for t in range(n_steps):
        obs, r, done, info = orc.step(rec["actions"][t][None])
        assert abs(r[0] - rec["reward"][t]) <= 1e-9, f"step {t}: reward {r[0]} vs {rec['reward'][t]}"
        assert bool(done[0]) == bool(rec["done"][t]), f"step {t}: terminated flag"
        got = [int(info["agent_kills"][0]), int(info["allies_kills"][0]), int(info["deads"][0]), int(info["current_wave"][0])]
        assert got == [int(v) for v in rec["info"][t]], f"step {t}: info {got} vs {rec['info'][t]}"
        _check_obs(rec, k, obs, orc, f"step {t}"); k += 1
        if done[0]:
            obs = orc.reset()
            _check_obs(rec, k, obs, orc, f"reset after step {t}"); k += 1
    assert [int(orc.spawn_ctr[0]), int(orc.hit_ctr[0]), int(orc.phys_ctr[0]), int(orc.obs_call[0])] == [int(v) for v in rec["counters"]]


def test_level5_golden_cases_exist():
    assert len(CASES) >= 4 and len(FUSION_CASES) >= 2
    for path in FUSION_CASES:                # the student stacks differ from the returned ones (own fusion draws)
        rec = load_recording(path)
        assert "stacked_student" in rec.files
        assert (rec["mask_student"] != rec["mask"]).any() and (rec["chosen_student"] != rec["chosen"]).any()


# recordings of Level5DumbMultiObs + Level5DumbMultiObjectTask (the data-collection env: every armed wingman observes)
DUMB_CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "l5dumb_*.npz")))


def replay_dumb(rec, make_oracle):
    from oracle.level5_oracle import LEVEL5_DUMB
    seed, env_index, n_steps, _ = (int(v) for v in rec["meta"])
    cfg = dataclasses.replace(LEVEL5_DUMB, noise_ratio=float(rec["noise_ratio"]), step_increment=int(rec["step_increment"]))
    orc = make_oracle(cfg, seed, env_index)
    assert int(orc.agent[0]) == int(rec["agent_slot"])
    L = cfg.n_lw

    def check(k, obs, tag):
        assert (rec["armed"][k] == orc.armed[0]).all(), f"{tag}: armed flags differ"
        assert np.abs(rec["pos"][k] - orc.pos[0]).max() <= 1e-9, f"{tag}: positions differ"
        assert (rec["ammo"][k] == orc.ammo[0, :L]).all(), f"{tag}: ammunition differs"
        pres = rec["present"][k]
        assert (pres == obs["present"][0]).all(), f"{tag}: observing wingmen differ"
        for P in np.nonzero(pres)[0]:
            t2 = f"{tag} observer {P}"
            assert np.abs(rec["inertial"][k][P].astype(np.float64) - obs["inertial_data"][0, P]).max() <= 1e-6, f"{t2}: inertial"
            assert np.abs(rec["teacher_actions"][k][P] - obs["last_action"][0, P]).max() <= 1e-6, f"{t2}: teacher action"
            if not rec["was_reset"][k]:
                assert (rec["chosen"][k][P] == orc.mo_chosen[0, P]).all(), f"{t2}: draws {orc.mo_chosen[0, P].tolist()} vs {rec['chosen'][k][P].tolist()}"
            assert (rec["mask"][k][P] == obs["validity_mask"][0, P]).all(), f"{t2}: validity mask"
            got, want = obs["stacked_spheres"][0, P], rec["stacked"][k][P]
            assert np.array_equal(got < 1, want < 1), f"{t2}: stack marks different cells"
            assert np.abs(got.astype(np.float64) - want.astype(np.float64)).max() <= 1e-6, f"{t2}: stack"

    obs = orc.reset()
    k = 0
    check(k, obs, "reset"); k += 1
    for t in range(n_steps):
        obs, r, done, info = orc.step(np.zeros((1, 4)))
        assert abs(r[0] - rec["reward"][t]) <= 1e-9, f"step {t}: reward {r[0]} vs {rec['reward'][t]}"
        assert bool(done[0]) == bool(rec["done"][t]), f"step {t}: terminated flag"
        got = [int(info["agent_kills"][0]), int(info["allies_kills"][0]), int(info["deads"][0]), int(info["current_wave"][0])]
        assert got == [int(v) for v in rec["info"][t]], f"step {t}: info {got} vs {rec['info'][t]}"
        check(k, obs, f"step {t}"); k += 1
        if done[0]:
            obs = orc.reset()
            check(k, obs, f"reset after step {t}"); k += 1
    assert [int(orc.spawn_ctr[0]), int(orc.hit_ctr[0]), int(orc.phys_ctr[0]), int(orc.obs_call[0])] == [int(v) for v in rec["counters"]]


@pytest.mark.parametrize("path", DUMB_CASES, ids=[os.path.basename(p)[:-4] for p in DUMB_CASES])
def test_level5_dumb_multiobs_oracle_matches_reference_recording(path):
    replay_dumb(load_recording(path), lambda cfg, seed, env_index: Level5Oracle(cfg, 1, seed=seed, env_offset=env_index))


def test_level5_dumb_golden_cases_cover_the_paths():
    assert len(DUMB_CASES) >= 2
    recs = [load_recording(p) for p in DUMB_CASES]
    assert any(int(r["agent_slot"]) != 0 for r in recs)                       # a non-zero agent slot
    assert any(r["done"].any() for r in recs)                                 # a termination + reset
    assert any((~r["present"]).any() for r in recs)                           # a disarmed wingman drops out of the lists
    assert all((r["mask"].sum(2)[r["present"]] >= 0).all() for r in recs)


# recordings of Level52BTEvaluationEnvironment + Level52BTEvaluationTask (two behaviour-tree wingmen, no observation)
EVAL_CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "l5eval2bt_*.npz")))


def replay_eval2bt(rec, make_oracle):
    from oracle.level5_oracle import LEVEL5_EVAL2BT
    seed, env_index, n_steps, _ = (int(v) for v in rec["meta"])
    cfg = dataclasses.replace(LEVEL5_EVAL2BT, noise_ratio=float(rec["noise_ratio"]), max_step=int(rec["max_step"]))
    orc = make_oracle(cfg, seed, env_index)

    def check(k, tag):
        assert (rec["armed"][k] == orc.armed[0]).all(), f"{tag}: armed flags differ"
        assert np.abs(rec["pos"][k] - orc.pos[0]).max() <= 1e-9, f"{tag}: positions differ"
        assert (rec["ammo"][k] == orc.ammo[0, :2]).all(), f"{tag}: ammunition differs"

    orc.reset()
    k = 0
    check(k, "reset"); k += 1
    for t in range(n_steps):
        obs, r, done, info = orc.step(np.zeros((1, 4)))
        assert r[0] == 0.0 and bool(done[0]) == bool(rec["done"][t]), f"step {t}: reward / terminated flag"
        # kills_per_drone: slot 0 counts as "the agent" of the shared engagement code, slot 1 as the ally
        got = [int(info["agent_kills"][0]), int(info["allies_kills"][0]), int(info["deads"][0]), int(info["current_wave"][0])]
        want = [int(rec["kills"][t][0]), int(rec["kills"][t][1]), int(rec["info"][t][0]), int(rec["info"][t][1])]
        assert got == want, f"step {t}: info {got} vs {want}"
        check(k, f"step {t}"); k += 1
        if done[0]:
            orc.reset()
            check(k, f"reset after step {t}"); k += 1
    assert [int(orc.spawn_ctr[0]), int(orc.hit_ctr[0]), int(orc.phys_ctr[0])] == [int(v) for v in rec["counters"]]


@pytest.mark.parametrize("path", EVAL_CASES, ids=[os.path.basename(p)[:-4] for p in EVAL_CASES])
def test_level5_eval2bt_oracle_matches_reference_recording(path):
    replay_eval2bt(load_recording(path), lambda cfg, seed, env_index: Level5Oracle(cfg, 1, seed=seed, env_offset=env_index))

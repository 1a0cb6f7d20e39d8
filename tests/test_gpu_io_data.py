"""collect_data (dronechase_b200/io_data.py) on a B200: (teacher observation, student observation, teacher action)
triples of Level5FusionEnvironment written from device batches -- what the parts hold equals what a second,
identically seeded env shows, the requested count is met exactly, the layout is the reference's
(src/core/rl_framework/utils/io_data.py:106-165)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _policy(seed):
    g = torch.Generator(device="cuda"); g.manual_seed(seed)

    def act(teacher_obs):
        E = teacher_obs["inertial_data"].shape[0]
        a = torch.rand(E, 4, generator=g, device="cuda") * 2 - 1
        a[:, 3] = a[:, 3].abs()
        return a
    return act


def test_collect_data_matches_a_replica_env(tmp_path):
    from dronechase_b200 import BatchedThreatEngageEnv
    from dronechase_b200.io_data import DatasetWriter, MultiFileDataset, _open_part, collect_data
    E, N = 192, 1500
    env = BatchedThreatEngageEnv("level5_fusion", n_envs=E, seed=11, auto_reset=True, with_student=True)
    with DatasetWriter(str(tmp_path), samples_per_file=400, backend="npz") as w:
        res = collect_data(env, _policy(3), w, max_observations_collected=N)
    assert res["observations"] == N
    ds = MultiFileDataset(str(tmp_path))
    assert len(ds) == N and len(ds.file_paths) == 4
    parts = [_open_part(p) for p in ds.file_paths]
    cat = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    assert cat["student/validity_mask"].any(axis=1).all()
    marked = (cat["student/stacked_spheres"] < 1).any(axis=(2, 3, 4))
    assert not (marked & ~cat["student/validity_mask"]).any() and marked.any()
    assert not cat["teacher/lidar"].any() and np.array_equal(cat["teacher/inertial_data"], cat["student/inertial_data"])
    # replica: same seed, same teacher -> the same rows in the same order
    rep = BatchedThreatEngageEnv("level5_fusion", n_envs=E, seed=11, auto_reset=True, with_student=True)
    pol = _policy(3)
    rep.reset()
    k = 0
    while k < N:
        tobs = {"inertial_data": rep.obs["inertial_data"], "last_action": rep.obs["last_action"]}
        a = pol(tobs)
        valid = rep.student_obs["validity_mask"].any(dim=1).cpu().numpy()
        rows = np.nonzero(valid)[0][:N - k]
        n = len(rows)
        if n:
            assert np.array_equal(cat["student/stacked_spheres"][k:k + n], rep.student_obs["stacked_spheres"].cpu().numpy()[rows])
            assert np.array_equal(cat["student/validity_mask"][k:k + n], rep.student_obs["validity_mask"].cpu().numpy()[rows])
            assert np.array_equal(cat["student/inertial_data"][k:k + n], rep.obs["inertial_data"].cpu().numpy()[rows])
            assert np.array_equal(cat["teacher_actions"][k:k + n], a.cpu().numpy()[rows])
        k += n
        rep.step(a)
    env.close(); rep.close()


def test_collect_data_against_the_oracle(tmp_path):
    """An INDEPENDENT source for the dataset: the float64 level5 oracle is stepped with the very actions the teacher took
    on the GPU; every row the writer stored -- student stack, validity mask, inertial vector, teacher action -- must be the
    oracle's ``student_observation`` of that env at that step, in env order (io_data.py:67-104 stores the pair BEFORE the
    step).  f64 build of the simulator, so marked cells and masks are exact."""
    from dronechase_b200 import BatchedThreatEngageEnv
    from dronechase_b200.io_data import DatasetWriter, MultiFileDataset, _open_part, collect_data
    from oracle.level5_oracle import LEVEL5_FUSION, Level5Oracle
    E, N, seed = 8, 600, 23
    env = BatchedThreatEngageEnv("level5_fusion", n_envs=E, seed=seed, auto_reset=True, precision="f64", with_student=True)
    taken = []
    g = torch.Generator(device="cuda"); g.manual_seed(7)

    def teacher(tobs):
        a = torch.rand(E, 4, generator=g, device="cuda") * 2 - 1
        a[:, 3] = a[:, 3].abs()
        taken.append(a.cpu().numpy())
        return a
    with DatasetWriter(str(tmp_path), samples_per_file=250, backend="npz") as w:
        res = collect_data(env, teacher, w, max_observations_collected=N)
    assert res["observations"] == N
    ds = MultiFileDataset(str(tmp_path))
    parts = [_open_part(p) for p in ds.file_paths]
    cat = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    assert len(cat["teacher_actions"]) == N

    orc = Level5Oracle(LEVEL5_FUSION, E, seed=seed, auto_reset=True)
    ref = orc.reset()
    k = marked = 0
    for t, a in enumerate(taken):
        rows = np.nonzero(ref["student_validity_mask"].any(axis=1))[0][:N - k]
        n = len(rows)
        if n:
            got, want = cat["student/stacked_spheres"][k:k + n], ref["student_stacked_spheres"][rows]
            assert np.array_equal(cat["student/validity_mask"][k:k + n], ref["student_validity_mask"][rows]), f"step {t}: mask"
            assert np.array_equal(got < 1, want < 1), f"step {t}: marked cells of the stored student stack"
            assert np.abs(got - want).max() < 1e-6, f"step {t}: stored student stack"
            assert np.abs(cat["student/inertial_data"][k:k + n] - ref["inertial_data"][rows]).max() < 1e-6, f"step {t}: inertial"
            assert np.array_equal(cat["teacher_actions"][k:k + n], a[rows]), f"step {t}: teacher action"
            marked += int((want < 1).sum())
        k += n
        if k >= N:
            break
        ref, _, _, _ = orc.step(a.astype(np.float64))
    assert k == N and marked > 500
    env.close()

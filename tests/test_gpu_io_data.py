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

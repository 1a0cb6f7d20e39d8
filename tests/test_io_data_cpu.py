"""Host logic of the teacher/student dataset writer (dronechase_b200/io_data.py) with CPU tensors: the layout of the
reference's IOData.save_to_hdf5 (src/core/rl_framework/utils/io_data.py:106-165), files of exactly N samples, row
selection by validity mask, order preserved, reader = MultiH5Dataset's contract (:13-52)."""
import numpy as np
import torch

from dronechase_b200.io_data import STUDENT_KEYS, TEACHER_KEYS, DatasetWriter, IOData, MultiFileDataset, _open_part


def _batch(rng, B, tag0):
    student = {"stacked_spheres": torch.from_numpy(rng.uniform(0, 1, (B, 6, 3, 13, 26)).astype(np.float32)),
               "validity_mask": torch.from_numpy(rng.rand(B, 6) < 0.3),
               "inertial_data": torch.from_numpy(rng.uniform(-1, 1, (B, 15)).astype(np.float32)),
               "last_action": torch.from_numpy(rng.uniform(-1, 1, (B, 4)).astype(np.float32))}
    teacher = {"lidar": torch.zeros(B, 2, 13, 26), "inertial_data": student["inertial_data"], "last_action": student["last_action"]}
    actions = torch.arange(tag0, tag0 + B, dtype=torch.float32)[:, None].repeat(1, 4)      # row tag: global sample number
    return teacher, student, actions


def test_writer_layout_and_order(tmp_path):
    rng = np.random.RandomState(0)
    folder = str(tmp_path / "collect_and_save")
    kept_tags, kept_inertial = [], []
    with DatasetWriter(folder, samples_per_file=100, backend="npz") as w:
        tag = 0
        for _ in range(9):
            B = int(rng.randint(20, 70))
            teacher, student, actions = _batch(rng, B, tag)
            n = w.append(teacher, student, actions)
            valid = student["validity_mask"].any(dim=1).numpy()
            assert n == int(valid.sum())
            kept_tags += list(np.arange(tag, tag + B)[valid]); kept_inertial.append(student["inertial_data"].numpy()[valid])
            tag += B
        assert w.samples_appended == len(kept_tags)
    ds = MultiFileDataset(folder)
    assert len(ds) == len(kept_tags)
    n_files = len(ds.file_paths)
    assert n_files == -(-len(kept_tags) // 100)
    inertial = np.concatenate(kept_inertial)
    # parts sort as io_data0, io_data1, ... (fewer than ten here): global order = append order
    for file_id, path in enumerate(ds.file_paths):
        part = _open_part(path)
        assert set(part) == {"teacher/" + k for k in TEACHER_KEYS} | {"student/" + k for k in STUDENT_KEYS} | {"teacher_actions"}
        rows = part["teacher_actions"].shape[0]
        assert rows == (100 if file_id < n_files - 1 else len(kept_tags) - 100 * (n_files - 1))
        assert part["student/validity_mask"].dtype == np.bool_ and part["student/validity_mask"].any(axis=1).all()
        assert part["student/stacked_spheres"].dtype == np.float32 and part["student/stacked_spheres"].shape == (rows, 6, 3, 13, 26)
        assert part["teacher/lidar"].shape == (rows, 2, 13, 26) and not part["teacher/lidar"].any()
        lo = 100 * file_id
        assert np.array_equal(part["teacher_actions"][:, 0], np.asarray(kept_tags[lo:lo + rows], dtype=np.float32))
        assert np.array_equal(part["student/inertial_data"], inertial[lo:lo + rows])
        assert np.array_equal(part["teacher/inertial_data"], inertial[lo:lo + rows])
    obs, target = ds[137]
    assert set(obs) == set(STUDENT_KEYS) and obs["validity_mask"].dtype == torch.bool
    assert float(target[0]) == float(kept_tags[137]) and target.shape == (4,)
    io = IOData(folder)
    ob, tg = next(iter(io.get_loader(batch_size=32, shuffle=False)))
    assert ob["stacked_spheres"].shape == (32, 6, 3, 13, 26) and tg.shape == (32, 4)
    folds = list(io.cross_validation_loaders(k_folds=3, batch_size=64))
    assert len(folds) == 3 and len(folds[0][0].dataset) + len(folds[0][1].dataset) == len(ds)


def test_writer_explicit_valid_and_empty_batches(tmp_path):
    rng = np.random.RandomState(1)
    folder = str(tmp_path / "d")
    w = DatasetWriter(folder, samples_per_file=50, backend="npz")
    teacher, student, actions = _batch(rng, 40, 0)
    assert w.append(teacher, student, actions, valid=torch.zeros(40, dtype=torch.bool)) == 0
    sel = torch.zeros(40, dtype=torch.bool); sel[[3, 7, 31]] = True
    assert w.append(None, student, actions, valid=sel) == 3          # student-only parts (collect_and_save.py:51-97)
    w.close()
    part = _open_part(MultiFileDataset(folder).file_paths[0])
    assert "teacher/lidar" not in part and list(part["teacher_actions"][:, 0]) == [3.0, 7.0, 31.0]
    assert np.array_equal(part["student/stacked_spheres"], student["stacked_spheres"].numpy()[[3, 7, 31]])


def test_second_collection_appends_after_the_first(tmp_path):
    """The reference opens its parts in 'a' mode (io_data.py:106-165): a second run into the same folder continues after
    the highest existing part instead of overwriting part 0; overwrite=True starts clean."""
    rng = np.random.RandomState(2)
    folder = str(tmp_path / "d")
    for run in range(2):
        with DatasetWriter(folder, samples_per_file=10, backend="npz") as w:
            teacher, student, actions = _batch(rng, 25, 100 * run)
            w.append(teacher, student, actions, valid=torch.ones(25, dtype=torch.bool))
    ds = MultiFileDataset(folder)
    assert len(ds.file_paths) == 6 and len(ds) == 50
    tags = sorted(float(ds[i][1][0]) for i in range(len(ds)))
    assert tags == [float(t) for t in list(range(25)) + list(range(100, 125))]
    with DatasetWriter(folder, samples_per_file=10, backend="npz", overwrite=True) as w:
        teacher, student, actions = _batch(rng, 5, 0)
        w.append(teacher, student, actions, valid=torch.ones(5, dtype=torch.bool))
    assert len(MultiFileDataset(folder)) == 5


def test_h5_backend_with_a_stand_in_h5py(tmp_path, monkeypatch):
    """h5py is not in this image, so the 'h5' branch (the reference's container, io_data.py:106-165) runs here against a
    minimal stand-in with h5py's File / create_dataset / group-by-path / [()] surface; wherever the real h5py is
    installed the same test uses it."""
    import pickle
    from dronechase_b200 import io_data

    class _DS:
        def __init__(self, a): self.a = np.asarray(a); self.shape = self.a.shape
        def __getitem__(self, k): return self.a[k]

    class _File(dict):
        def __init__(self, path, mode):
            super().__init__(); self.path, self.mode, self.flat = path, mode, {}
            if mode == "r":
                with open(path, "rb") as fh:
                    for name, a in pickle.load(fh).items():
                        g, parts = self, name.split("/")
                        for q in parts[:-1]:
                            g = g.setdefault(q, {})
                        g[parts[-1]] = _DS(a)
        def create_dataset(self, name, data=None, **kw):
            assert kw["maxshape"][0] is None and kw["maxshape"][1:] == data.shape[1:] and kw["chunks"] is True
            self.flat[name] = np.asarray(data)
        def __enter__(self): return self
        def __exit__(self, *exc):
            if self.mode != "r":
                with open(self.path, "wb") as fh:
                    pickle.dump(self.flat, fh)

    if io_data.h5py is None:
        monkeypatch.setattr(io_data, "h5py", type("h5py", (), {"File": _File}))
    rng = np.random.RandomState(3)
    folder = str(tmp_path / "h5")
    with DatasetWriter(folder, samples_per_file=8, backend="h5") as w:
        teacher, student, actions = _batch(rng, 20, 0)
        w.append(teacher, student, actions, valid=torch.ones(20, dtype=torch.bool))
    ds = MultiFileDataset(folder)
    assert [p.rsplit(".", 1)[1] for p in ds.file_paths] == ["h5"] * 3 and len(ds) == 20
    part = _open_part(ds.file_paths[0])
    assert set(part) == {"teacher/" + k for k in TEACHER_KEYS} | {"student/" + k for k in STUDENT_KEYS} | {"teacher_actions"}
    assert np.array_equal(part["student/inertial_data"], student["inertial_data"].numpy()[:8])
    obs, target = ds[19]
    assert float(target[0]) == 19.0 and obs["stacked_spheres"].shape == (6, 3, 13, 26)

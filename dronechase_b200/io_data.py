"""Teacher/student dataset writer and reader.

The reference's threatsense pipeline collects (student observation, teacher action) pairs one sample at a time:
``IOData.collect_data`` steps one env and calls ``save_to_hdf5`` per sample, which re-opens the file and grows every
dataset by one row (src/core/rl_framework/utils/io_data.py:67-165; the threaded variant of
apps/threatsense_runner/collect_and_save.py:51-97 does the same per 1000-sample part).  Here a whole device batch is
appended per env step:

    env = BatchedThreatEngageEnv("level5_fusion", n_envs=4096, with_student=True)
    w = DatasetWriter("out/collect_and_save")
    collect_data(env, teacher_policy, w, max_observations_collected=1_000_000)

Layout = the reference's (io_data.py:106-165): groups ``teacher/{lidar, inertial_data, last_action}`` and
``student/{stacked_spheres, validity_mask, inertial_data, last_action}`` (float32, the mask bool) plus the dataset
``teacher_actions``; ``samples_per_file`` rows per file (1000: io_data.py:83), files ``io_data<k>``.  Backend: ``.h5``
when h5py is importable, else ``.npz`` holding the same keys as ``"group/name"`` (h5py is not in this image); the
reader (``MultiFileDataset`` = ``MultiH5Dataset``, io_data.py:13-52) takes both.

Data path: rows are selected on the device (``valid`` mask: the reference drops samples whose validity mask is all
False, collect_and_save.py:100-110), copied in one D2H per key into pinned staging, cut into files of exactly
``samples_per_file`` rows and written by a background thread, so the env loop never waits for the disk.
"""
from __future__ import annotations

import glob
import os
import re
import queue
import threading
import time
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch

try:                                    # optional: the reference's container format
    import h5py                         # type: ignore
except Exception:                       # noqa: BLE001 -- absent in this image
    h5py = None

TEACHER_KEYS = ("lidar", "inertial_data", "last_action")                                     # level5_envrionment.py:336-340
STUDENT_KEYS = ("stacked_spheres", "validity_mask", "inertial_data", "last_action")          # :342-346


def _np_dtype(key: str):
    return np.bool_ if key == "validity_mask" else np.float32        # io_data.py:131-135


class DatasetWriter:
    def __init__(self, folder_path: str, samples_per_file: int = 1000, backend: Optional[str] = None,
                 file_stem: str = "io_data", max_pending_files: int = 8, disk_threads: int = 4, overwrite: bool = False):
        if backend is None:
            backend = "h5" if h5py is not None else "npz"
        if backend == "h5" and h5py is None:
            raise RuntimeError("backend 'h5' needs h5py, which is not installed; use backend='npz'")
        if backend not in ("h5", "npz"):
            raise ValueError(f"unknown backend {backend!r}")
        self.folder_path, self.backend, self.file_stem = folder_path, backend, file_stem
        self.samples_per_file = int(samples_per_file)
        os.makedirs(folder_path, exist_ok=True)
        # The reference opens its parts in 'a' mode and appends (io_data.py:106-165): a second collection into the same
        # folder must not overwrite the low-numbered parts of the first.  Continue after the highest existing part
        # (``overwrite=True`` removes the previous run's parts instead, all of them).
        self.file_id = 0
        pat = re.compile(re.escape(file_stem) + r"(\d+)\.(npz|h5)$")
        for name in os.listdir(folder_path):
            m = pat.match(name)
            if not m:
                continue
            if overwrite:
                os.remove(os.path.join(folder_path, name))
            else:
                self.file_id = max(self.file_id, int(m.group(1)) + 1)
        self.samples_written = 0                      # rows handed to the disk thread (complete files only)
        self._pending: Dict[str, List[np.ndarray]] = {}
        self._pending_rows = 0
        self._pin: Dict[str, torch.Tensor] = {}
        self._q: "queue.Queue" = queue.Queue(maxsize=max_pending_files)
        self._error: Optional[BaseException] = None
        # parts are independent files: several writers (the reference's threaded collector uses 4,
        # collect_and_save.py:153); np.savez spends its time in the zip CRC, which releases the GIL
        self._threads = [threading.Thread(target=self._disk_loop, name=f"dronechase-dataset-writer-{i}", daemon=True)
                         for i in range(max(1, int(disk_threads)))]
        for t in self._threads:
            t.start()

    def _pinned(self, name: str, like: torch.Tensor) -> torch.Tensor:
        """A pinned landing zone for `like` (cudaHostAlloc costs milliseconds: allocate once per key, grow by doubling)."""
        n = like.shape[0]
        buf = self._pin.get(name)
        if buf is None or buf.shape[0] < n or buf.shape[1:] != like.shape[1:] or buf.dtype != like.dtype:
            cap = max(n, 2 * buf.shape[0] if buf is not None and buf.shape[1:] == like.shape[1:] else n)
            buf = torch.empty((cap,) + tuple(like.shape[1:]), dtype=like.dtype, pin_memory=True)
            self._pin[name] = buf
        return buf[:n]

    # ------------------------------------------------------------------ append
    @property
    def samples_appended(self) -> int:
        return self.samples_written + self._pending_rows

    def append(self, teacher_obs: Optional[Dict[str, torch.Tensor]], student_obs: Dict[str, torch.Tensor],
               teacher_actions: torch.Tensor, valid: Optional[torch.Tensor] = None) -> int:
        """Append one batch ([B, ...] tensors, device or host).  ``valid`` ([B] bool) selects rows; default = the rows
        whose student validity mask has at least one True.  Returns the number of rows appended."""
        if self._error is not None:
            raise RuntimeError("dataset writer thread failed") from self._error
        mask = student_obs["validity_mask"]
        if valid is None:
            valid = mask.reshape(mask.shape[0], -1).any(dim=1)
        idx = torch.nonzero(valid.reshape(-1), as_tuple=False).reshape(-1)
        n = int(idx.numel())                          # one small D2H sync per batch
        if n == 0:
            return 0
        cols: Dict[str, np.ndarray] = {}
        staged: List[Tuple[str, torch.Tensor]] = []

        def stage(name: str, t: torch.Tensor, key: str):
            rows = t.index_select(0, idx)
            want = torch.bool if key == "validity_mask" else torch.float32
            if rows.dtype != want:
                rows = rows.to(want)
            if rows.is_cuda:                          # one D2H per key into pinned memory, all in flight together
                host = self._pinned(name, rows)
                host.copy_(rows, non_blocking=True)
                rows = host
            staged.append((name, rows))

        if teacher_obs is not None:
            for k in TEACHER_KEYS:
                stage("teacher/" + k, teacher_obs[k], k)
        for k in STUDENT_KEYS:
            stage("student/" + k, student_obs[k], k)
        stage("teacher_actions", teacher_actions, "teacher_actions")
        from_device = any(t.is_pinned() for _, t in staged) if torch.cuda.is_available() else False
        if from_device:
            torch.cuda.current_stream().synchronize()
        for name, t in staged:                        # the pinned landing zones are reused by the next append: copy out
            cols[name] = t.numpy().copy() if from_device else t.numpy()
        for name, a in cols.items():
            self._pending.setdefault(name, []).append(a)
        self._pending_rows += n
        while self._pending_rows >= self.samples_per_file:
            self._cut(self.samples_per_file)
        return n

    def _cut(self, rows: int):
        """Hand the first `rows` pending rows to the disk thread as one file."""
        out = {}
        for name, parts in self._pending.items():
            cat = parts[0] if len(parts) == 1 else np.concatenate(parts, axis=0)
            out[name] = np.ascontiguousarray(cat[:rows])
            rest = cat[rows:]
            self._pending[name] = [rest] if len(rest) else []
        self._pending_rows -= rows
        self._q.put((self.file_id, out))              # blocks when the disk is max_pending_files behind
        self.file_id += 1
        self.samples_written += rows

    # -------------------------------------------------------------------- disk
    def _path(self, file_id: int) -> str:
        return os.path.join(self.folder_path, f"{self.file_stem}{file_id}.{self.backend}")

    def _disk_loop(self):
        while True:
            item = self._q.get()
            if item is None:
                return
            try:
                file_id, cols = item
                path, tmp = self._path(file_id), self._path(file_id) + ".tmp"
                if self.backend == "h5":
                    with h5py.File(tmp, "w") as f:
                        for name, a in cols.items():
                            f.create_dataset(name, data=a, maxshape=(None,) + a.shape[1:], chunks=True)
                else:
                    with open(tmp, "wb") as fh:
                        np.savez(fh, **cols)
                os.replace(tmp, path)                 # a reader never sees a half-written part
            except BaseException as e:                # noqa: BLE001 -- surfaced by the next append / close
                self._error = e
            finally:
                self._q.task_done()

    def close(self, flush_partial: bool = True):
        """Write the remaining rows (a last, shorter file) and wait for the disk thread."""
        if not self._threads:
            return
        if flush_partial and self._pending_rows > 0:
            self._cut(self._pending_rows)
        for _ in self._threads:
            self._q.put(None)
        for t in self._threads:
            t.join()
        self._threads = []
        if self._error is not None:
            raise RuntimeError("dataset writer thread failed") from self._error

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def _part_rows(path: str) -> int:
    """Rows of one part without decompressing its observations (np.load indexes the zip lazily)."""
    if path.endswith(".npz"):
        with np.load(path) as z:
            return int(z["teacher_actions"].shape[0])
    if h5py is None:
        raise RuntimeError(f"{path}: reading .h5 parts needs h5py")
    with h5py.File(path, "r") as f:
        return int(f["teacher_actions"].shape[0])


def _open_part(path: str) -> Dict[str, np.ndarray]:
    if path.endswith(".npz"):
        with np.load(path) as z:
            return {k: z[k] for k in z.files}
    if h5py is None:
        raise RuntimeError(f"{path}: reading .h5 parts needs h5py")
    out = {}
    with h5py.File(path, "r") as f:
        for g in ("teacher", "student"):
            if g in f:
                for k in f[g]:
                    out[f"{g}/{k}"] = f[g][k][()]
        out["teacher_actions"] = f["teacher_actions"][()]
    return out


class MultiFileDataset(torch.utils.data.Dataset):
    """``MultiH5Dataset`` (io_data.py:13-52): every part of a folder as one dataset of
    (student observation dict, teacher action).  Parts are cached whole after the first touch (a part is 24 MB)."""

    def __init__(self, folder: str, pattern: str = "*"):
        paths = sorted(p for p in glob.glob(os.path.join(folder, pattern)) if p.endswith((".npz", ".h5")))
        if not paths:
            raise FileNotFoundError(f"no dataset parts in {folder} matching {pattern}")
        self.file_paths = paths
        self.index_map: List[Tuple[int, int]] = []
        for file_id, path in enumerate(paths):
            n = _part_rows(path)
            self.index_map.extend((file_id, i) for i in range(n))
        self._cache: Dict[int, Dict[str, np.ndarray]] = {}

    def __len__(self) -> int:
        return len(self.index_map)

    def _part(self, file_id: int):
        if file_id not in self._cache:
            if len(self._cache) >= 16:
                self._cache.pop(next(iter(self._cache)))
            self._cache[file_id] = _open_part(self.file_paths[file_id])
        return self._cache[file_id]

    def __getitem__(self, idx: int):
        file_id, i = self.index_map[idx]
        part = self._part(file_id)
        obs = {k: torch.as_tensor(part["student/" + k][i], dtype=torch.bool if k == "validity_mask" else torch.float32)
               for k in STUDENT_KEYS}
        return obs, torch.as_tensor(part["teacher_actions"][i], dtype=torch.float32)


def collect_data(env, teacher_policy: Callable[[Dict[str, torch.Tensor]], torch.Tensor], writer: DatasetWriter,
                 max_observations_collected: int = 5_000_000, log_every_s: float = 0.0) -> Dict[str, float]:
    """``IOData.collect_data`` (io_data.py:67-104) over a batched env: every step, the teacher acts on
    ``info["teacher_observation"]`` (lidar zeros(2,13,26), inertial_data, last_action) of every env, the env steps, and
    (teacher observation, ``info["student_observation"]``, teacher action) of the envs with a usable student stack go
    to ``writer``.  The pair stored is (observation the teacher saw, action it took), taken BEFORE the step.
    ``env``: a ``BatchedThreatEngageEnv`` created with ``with_student=True`` and ``auto_reset=True``."""
    if getattr(env, "student_obs", None) is None:
        raise ValueError("collect_data needs an env created with with_student=True (info['student_observation'])")
    E = env.n_envs
    dev = env.device
    teacher_lidar = torch.zeros(E, 2, 13, 26, dtype=torch.float32, device=dev)      # level5_envrionment.py:326
    env.reset()
    collected, steps, t0, t_log = 0, 0, time.time(), time.time()
    while collected < max_observations_collected:
        teacher_obs = {"lidar": teacher_lidar, "inertial_data": env.obs["inertial_data"], "last_action": env.obs["last_action"]}
        actions = teacher_policy(teacher_obs)
        room = max_observations_collected - collected
        valid = env.student_obs["validity_mask"].any(dim=1)
        if room < E:                                   # never overshoot the requested count
            valid = valid & (torch.cumsum(valid.to(torch.int64), 0) <= room)
        collected += writer.append(teacher_obs, env.student_obs, actions, valid)
        env.step(actions)
        steps += 1
        if log_every_s and time.time() - t_log > log_every_s:
            t_log = time.time()
            print(f"[INFO] Collected {collected} / {max_observations_collected}, "
                  f"Avg speed: {collected / (t_log - t0):.2f} obs/sec", flush=True)
    dt = time.time() - t0
    return {"observations": collected, "env_steps": steps * E, "seconds": dt, "obs_per_sec": collected / max(dt, 1e-9)}


def collect_data_multiobs(env, writer: DatasetWriter, max_observations_collected: int = 1_000_000,
                          log_every_s: float = 0.0) -> Dict[str, float]:
    """The loop of apps/threatsense_runner/collect_and_save.py:131-205 over the batched ``Level5DumbMultiObs``
    (preset ``level5_dumb_multiobs``): every env step, the student observation of every armed wingman whose validity mask
    has a True (``drop_invalid_student_obs`` :100-110) goes to ``writer`` with that wingman's behaviour-tree command as the
    teacher action; parts hold the ``student`` group and ``teacher_actions`` only (``save_to_hdf5`` :51-97)."""
    mo = getattr(env, "multi_obs", None)
    if mo is None:
        raise ValueError("collect_data_multiobs needs a level5_multi_obs env (preset 'level5_dumb_multiobs')")
    E, L = mo["present"].shape
    flat = {k: mo[k].view(E * L, *mo[k].shape[2:]) for k in STUDENT_KEYS}
    env.reset()
    collected, steps, t0, t_log = 0, 0, time.time(), time.time()
    while collected < max_observations_collected:
        env.step(None)
        steps += 1
        valid = (mo["present"] & mo["validity_mask"].any(dim=2)).reshape(-1)
        room = max_observations_collected - collected
        if room < E * L:
            valid = valid & (torch.cumsum(valid.to(torch.int64), 0) <= room)
        collected += writer.append(None, flat, flat["last_action"], valid)
        if log_every_s and time.time() - t_log > log_every_s:
            t_log = time.time()
            print(f"[INFO] Collected {collected} / {max_observations_collected} observations, "
                  f"Avg speed: {collected / (t_log - t0):.2f} obs/sec", flush=True)
    dt = time.time() - t0
    return {"observations": collected, "env_steps": steps * E, "seconds": dt, "obs_per_sec": collected / max(dt, 1e-9)}


class IOData:
    """``IOData`` (io_data.py:55-65,167-222): folder + dataset + loaders."""

    def __init__(self, folder_path: str):
        self.folder_path = folder_path
        os.makedirs(folder_path, exist_ok=True)
        self.dataset = MultiFileDataset(folder_path)

    def get_loader(self, batch_size: int = 1024, shuffle: bool = True, num_workers: int = 0):
        return torch.utils.data.DataLoader(self.dataset, batch_size=batch_size, shuffle=shuffle, num_workers=num_workers)

    def get_loader_limit(self, limit_size: int = 1_000_000, batch_size: int = 1024, shuffle: bool = True, rng=None,
                         num_workers: int = 0):
        rng = np.random.default_rng(42) if rng is None else rng
        indices = rng.choice(len(self.dataset), size=min(limit_size, len(self.dataset)), replace=False)
        return torch.utils.data.DataLoader(torch.utils.data.Subset(self.dataset, indices.tolist()), batch_size=batch_size,
                                           shuffle=shuffle, num_workers=num_workers)

    def cross_validation_loaders(self, k_folds: int = 5, batch_size: int = 256):
        n = len(self.dataset)
        indices = np.arange(n)
        np.random.shuffle(indices)
        fold = n // k_folds
        for f in range(k_folds):
            test_idx = indices[f * fold:(f + 1) * fold]
            train_idx = np.setdiff1d(indices, test_idx)
            yield (torch.utils.data.DataLoader(torch.utils.data.Subset(self.dataset, train_idx.tolist()), batch_size=batch_size, shuffle=True),
                   torch.utils.data.DataLoader(torch.utils.data.Subset(self.dataset, test_idx.tolist()), batch_size=batch_size, shuffle=True))

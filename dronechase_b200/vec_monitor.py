"""``VecMonitor`` for hosts without stable-baselines3, and the helper behind ``compat/``'s shadowing modules.

``ReinforcementLearningPipeline.create_vectorized_environment`` (src/core/rl_framework/utils/pipeline.py:32-61) returns
``VecMonitor(SubprocVecEnv(...))``; SB3's ``VecMonitor`` records per-env episode return / length / wall time and adds
``info["episode"] = {"r", "l", "t"}`` to the info dict of every env that finished.  ``VecMonitor`` below honours that
contract (and SB3's ``VecEnvWrapper`` surface: ``venv``, ``num_envs``, spaces, ``step_async/step_wait/step/reset/close``,
attribute forwarding) for the batched env, without turning the lazy ``InfoList`` into 65,536 dicts per step.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import time
from typing import Optional

import numpy as np


class VecMonitor:
    def __init__(self, venv, filename: Optional[str] = None, info_keywords=()):
        self.venv = venv
        self.num_envs = venv.num_envs
        self.observation_space, self.action_space = venv.observation_space, venv.action_space
        self.info_keywords = tuple(info_keywords)
        self.episode_count = 0
        self.t_start = time.time()
        self.episode_returns = np.zeros(self.num_envs, dtype=np.float32)
        self.episode_lengths = np.zeros(self.num_envs, dtype=np.int32)
        self._csv = None
        if filename is not None:                      # monitor.csv, SB3's ResultsWriter layout
            path = filename if filename.endswith("monitor.csv") else (
                os.path.join(filename, "monitor.csv") if os.path.isdir(filename) else filename + ".monitor.csv")
            self._csv = open(path, "wt")
            self._csv.write('#{"t_start": %r, "env_id": "dronechase_b200"}\n' % self.t_start)
            self._csv.write(",".join(("r", "l", "t") + self.info_keywords) + "\n")

    def reset(self):
        obs = self.venv.reset()
        self.episode_returns[:] = 0
        self.episode_lengths[:] = 0
        return obs

    def step_async(self, actions):
        self.venv.step_async(actions)

    def step_wait(self):
        obs, rewards, dones, infos = self.venv.step_wait()
        self.episode_returns += rewards
        self.episode_lengths += 1
        finished = np.nonzero(dones)[0]
        if len(finished):
            infos = _EpisodeInfos(infos)
            now = round(time.time() - self.t_start, 6)
            for i in finished:
                ep = {"r": float(self.episode_returns[i]), "l": int(self.episode_lengths[i]), "t": now}
                info = dict(infos[int(i)])
                for k in self.info_keywords:
                    ep[k] = info[k]
                info["episode"] = ep
                infos.override[int(i)] = info
                if self._csv is not None:
                    self._csv.write(",".join(str(ep[k]) for k in ("r", "l", "t") + self.info_keywords) + "\n")
            self.episode_count += len(finished)
            self.episode_returns[finished] = 0
            self.episode_lengths[finished] = 0
            if self._csv is not None:
                self._csv.flush()
        return obs, rewards, dones, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        if self._csv is not None:
            self._csv.close()
            self._csv = None
        return self.venv.close()

    def __getattr__(self, name):                      # VecEnvWrapper forwards unknown attributes to the wrapped env
        if name == "venv":
            raise AttributeError(name)
        return getattr(self.venv, name)


class _EpisodeInfos:
    """The wrapped env's info sequence with the dicts of the finished envs replaced."""

    def __init__(self, infos):
        self._infos, self.override = infos, {}

    def __len__(self):
        return len(self._infos)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        return self.override.get(i) or self._infos[i]

    def __iter__(self):
        return (self[i] for i in range(len(self)))


def vec_monitor_class():
    """SB3's VecMonitor when stable-baselines3 is importable (the contract the apps were written against), else ours."""
    try:
        from stable_baselines3.common.vec_env import VecMonitor as sb3_monitor
        return sb3_monitor
    except Exception:                                 # noqa: BLE001 -- absent in this image
        return VecMonitor


def load_shadowed(module_name: str, this_file: str):
    """For a ``compat/`` module that shadows a reference module of the same dotted name: find the NEXT file of that name
    along the parent package's ``__path__`` (the reference's own, when its ``src/`` is on ``sys.path`` behind
    ``compat/``), load it under ``<module_name>__reference`` and return it; ``None`` when there is none or it cannot be
    imported (e.g. stable-baselines3 missing)."""
    parent_name, _, leaf = module_name.rpartition(".")
    parent = sys.modules.get(parent_name)
    here = os.path.realpath(this_file)
    for d in list(getattr(parent, "__path__", [])):
        cand = os.path.join(d, leaf + ".py")
        if os.path.isfile(cand) and os.path.realpath(cand) != here:
            alias = f"{parent_name}.{leaf}__reference"
            if alias in sys.modules:
                return sys.modules[alias]
            spec = importlib.util.spec_from_file_location(alias, cand)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[alias] = mod
            try:
                spec.loader.exec_module(mod)
            except Exception:                         # noqa: BLE001
                sys.modules.pop(alias, None)
                return None
            return mod
    return None

"""The reference's module path src/core/rl_framework/utils/pipeline.py with the vectorisation boundary on the GPU.

``ReinforcementLearningPipeline.create_vectorized_environment`` (pipeline.py:32-61) and its multi-agent variants
(:64-119) build ``VecMonitor(SubprocVecEnv([lambda: environment(**kwargs)] * n_envs))`` -- one OS process per env.  Here
the same call returns ``VecMonitor(DroneChaseVecEnv(preset, n_envs))``: ONE device-resident batch behind the same SB3
``VecEnv`` contract, whenever ``environment`` is one of the dronechase_b200 facades (they carry a ``PRESET``).  Every other
member (``create_callback_list``, ``create_model``, ``evaluate``, ``save_*`` ...) is the reference's own, inherited from
its module when the reference's ``src/`` is on ``sys.path`` behind ``compat/`` and stable-baselines3 is installed.
``callbacklist`` / ``CallbackType`` are re-exported like the reference module does (the apps import them from here).

Batch size: the apps call without ``n_envs`` (the reference default is ``os.cpu_count()``); set the environment variable
``DRONECHASE_B200_ENVS`` (e.g. 65536) to choose the GPU batch, or pass ``n_envs``.
"""
from __future__ import annotations

import os

from dronechase_b200.vec_monitor import load_shadowed, vec_monitor_class

_ref = load_shadowed(__name__, __file__)
_Base = getattr(_ref, "ReinforcementLearningPipeline", object)
if _ref is not None:
    for _name in ("callbacklist", "CallbackType"):
        if hasattr(_ref, _name):
            globals()[_name] = getattr(_ref, _name)


def _default_envs() -> int:
    return int(os.environ.get("DRONECHASE_B200_ENVS", os.cpu_count() or 1))


def _gpu_vec_env(environment, env_kwargs, n_envs, env_args):
    from dronechase_b200 import preset
    from dronechase_b200.vec_env import DroneChaseVecEnv
    valid = {k: v for k, v in env_kwargs.items() if k in env_args}          # pipeline.py:45-51
    over = {}
    if "dome_radius" in valid:
        over["dome_radius"] = float(valid["dome_radius"])
    if "rl_frequency" in valid:
        over["rl_frequency"] = int(valid["rl_frequency"])
    cfg = preset(environment.PRESET, **over)
    venv = DroneChaseVecEnv(cfg, n_envs=n_envs, seed=int(os.environ.get("DRONECHASE_B200_SEED", "0")),
                            device=int(os.environ.get("LOCAL_RANK", "0")))
    return vec_monitor_class()(venv)


class ReinforcementLearningPipeline(_Base):
    @staticmethod
    def create_vectorized_environment(environment, env_kwargs: dict = {}, n_envs: int | None = None, GUI=False,
                                      env_args=None):
        env_kwargs = dict(env_kwargs)
        env_kwargs["GUI"] = GUI
        env_args = ["dome_radius", "rl_frequency", "GUI"] if env_args is None else env_args
        if hasattr(environment, "PRESET") and not GUI:
            return _gpu_vec_env(environment, env_kwargs, _default_envs() if n_envs is None else int(n_envs), env_args)
        if _Base is object:
            raise RuntimeError(f"{getattr(environment, '__name__', environment)!r} is not a dronechase_b200 environment and the "
                               "reference's own pipeline (stable-baselines3 SubprocVecEnv) is not importable here")
        return _Base.create_vectorized_environment(environment, env_kwargs, n_envs or (os.cpu_count() or 1), GUI, env_args)

    @staticmethod
    def create_vectorized_multi_agent_environment(environment, env_kwargs: dict, n_envs: int | None = None, GUI=False):
        # pipeline.py:64-92: same, env_args + "model_path" (exp05's second policy), the reference caps n_envs at 4
        # "because 16 is too much for the cpu" -- no such cap for one GPU batch
        env_kwargs = dict(env_kwargs)
        env_kwargs["GUI"] = GUI
        if hasattr(environment, "PRESET") and not GUI:
            return _gpu_vec_env(environment, env_kwargs, _default_envs() if n_envs is None else int(n_envs),
                                ["dome_radius", "rl_frequency", "model_path", "GUI"])
        if _Base is object:
            raise RuntimeError("not a dronechase_b200 environment and the reference's own pipeline is not importable here")
        return _Base.create_vectorized_multi_agent_environment(environment, env_kwargs, n_envs or (os.cpu_count() or 1), GUI)

    @staticmethod
    def create_vectorized_multi_agent_v2_environment(environment, env_kwargs: dict, n_envs: int | None = None, GUI=False):
        # pipeline.py:94-119: the DummyVecEnv variant
        env_kwargs = dict(env_kwargs)
        env_kwargs["GUI"] = GUI
        if hasattr(environment, "PRESET") and not GUI:
            return _gpu_vec_env(environment, env_kwargs, _default_envs() if n_envs is None else int(n_envs),
                                ["dome_radius", "rl_frequency", "GUI"])
        if _Base is object:
            raise RuntimeError("not a dronechase_b200 environment and the reference's own pipeline is not importable here")
        return _Base.create_vectorized_multi_agent_v2_environment(environment, env_kwargs, n_envs or (os.cpu_count() or 1), GUI)

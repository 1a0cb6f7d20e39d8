"""Test helper: a meta-path finder that serves empty stand-ins for third-party packages the reference imports but this
image lacks (stable_baselines3, optuna, gymnasium, pybullet, PyFlyt, pynput ...).  Any attribute of a stand-in module is
a dummy class created on demand, so ``from stable_baselines3.common.vec_env import SubprocVecEnv, VecMonitor`` works;
``VecMonitor`` / ``SubprocVecEnv`` / ``DummyVecEnv`` record what they were given."""
import importlib.abc
import importlib.machinery
import sys
import types

STUBBED = ("stable_baselines3", "optuna", "gymnasium", "pybullet", "pybullet_data", "pybullet_utils", "PyFlyt", "pynput",
           "h5py", "tensorboard", "sb3_contrib", "openpyxl", "torch_geometric", "pytorch3d")


class _Meta(type):
    def __getattr__(cls, name):                      # gym.spaces.Dict, Key.up ...: nested dummies on demand
        if name.startswith("__"):
            raise AttributeError(name)
        sub = _Meta(name, (_Recorder,), {"__module__": cls.__module__})
        setattr(cls, name, sub)
        return sub


class _Recorder(metaclass=_Meta):
    def __init__(self, *args, **kwargs):
        self.args, self.kwargs = args, kwargs
        if args and hasattr(args[0], "num_envs"):
            self.venv = args[0]
            self.num_envs = args[0].num_envs


class _StubModule(types.ModuleType):
    __path__ = []                                    # a package: submodules resolve through the finder

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = _Meta(name, (_Recorder,), {"__module__": self.__name__})
        setattr(self, name, cls)
        return cls


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in STUBBED:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


def install():
    if not any(isinstance(f, _Finder) for f in sys.meta_path):
        sys.meta_path.insert(0, _Finder())

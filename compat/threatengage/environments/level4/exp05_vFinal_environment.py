"""Drop-in for src/threatengage/environments/level4/exp05_vFinal_environment.py (single-env view of the GPU batch)."""
from dronechase_b200.gym_env import Exp05vFinalEnvironment  # noqa: F401

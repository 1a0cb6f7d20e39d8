"""Drop-in for src/threatengage/environments/level4/exp03_vFinal_environment.py (single-env view of the GPU batch)."""
from dronechase_b200.gym_env import Exp03vFinalEnvironment  # noqa: F401

"""Drop-in for src/threatengage/environments/level4/evaluation_environment.py (single-env view of the GPU batch)."""
from dronechase_b200.gym_env import EvaluationEnvironment  # noqa: F401

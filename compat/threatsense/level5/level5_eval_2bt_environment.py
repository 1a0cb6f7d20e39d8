"""Drop-in for src/threatsense/level5/level5_eval_2bt_environment.py (single-env view of the GPU batch)."""
from dronechase_b200.gym_env import Level52BTEvaluationEnvironment  # noqa: F401

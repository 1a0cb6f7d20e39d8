"""Drop-in for src/threatsense/level5/level5_dumb_multiobs.py (single-env view of the GPU batch)."""
from dronechase_b200.gym_env import Level5DumbMultiObs  # noqa: F401

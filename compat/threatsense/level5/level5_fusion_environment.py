"""Drop-in for src/threatsense/level5/level5_fusion_environment.py (single-env view of the GPU batch)."""
from dronechase_b200.gym_env import Level5FusionEnvironment  # noqa: F401

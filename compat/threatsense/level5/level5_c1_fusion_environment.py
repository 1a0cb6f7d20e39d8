"""Drop-in for src/threatsense/level5/level5_c1_fusion_environment.py (single-env view of the GPU batch)."""
from dronechase_b200.gym_env import Level5C1FusionEnvironment  # noqa: F401

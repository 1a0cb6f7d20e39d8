"""Drop-in for src/threatengage/environments/level3/pyflyt_level3_environment_v2.py."""
from dronechase_b200.gym_env import PyflytL3EnviromentV2  # noqa: F401

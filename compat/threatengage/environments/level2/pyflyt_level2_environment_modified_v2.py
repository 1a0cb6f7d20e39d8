"""Drop-in for src/threatengage/environments/level2/pyflyt_level2_environment_modified_v2.py."""
from dronechase_b200.gym_env import PyflytL2EnviromentModifiedV2  # noqa: F401

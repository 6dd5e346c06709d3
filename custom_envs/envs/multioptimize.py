from custom_envs_b200.envs.multioptimize import *  # noqa: F401,F403
from custom_envs_b200.envs.multioptimize import MultiOptimize, VersionType  # noqa: F401

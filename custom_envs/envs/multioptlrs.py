from custom_envs_b200.envs.multioptlrs import *  # noqa: F401,F403
from custom_envs_b200.envs.multioptlrs import MultiOptLRs, BOUNDS  # noqa: F401

from custom_envs_b200.envs import *  # noqa: F401,F403
from custom_envs_b200.envs import SINGLE_AGENT_ENVIRONMENTS  # noqa: F401

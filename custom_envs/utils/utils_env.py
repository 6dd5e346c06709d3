from custom_envs_b200.utils.utils_env import *  # noqa: F401,F403

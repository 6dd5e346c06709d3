from custom_envs_b200.wrappers import *  # noqa: F401,F403

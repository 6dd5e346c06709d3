from custom_envs_b200.dataset import *  # noqa: F401,F403

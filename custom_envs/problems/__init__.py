from custom_envs_b200.problems import *  # noqa: F401,F403

from custom_envs_b200.vectorize import *  # noqa: F401,F403

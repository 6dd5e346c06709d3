from custom_envs_b200.utils.utils_math import *  # noqa: F401,F403
from custom_envs_b200.utils.utils_math import use_random_state, normalize  # noqa: F401

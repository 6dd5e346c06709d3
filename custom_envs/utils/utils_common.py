from custom_envs_b200.utils.utils_common import *  # noqa: F401,F403
from custom_envs_b200.utils.utils_common import History, shuffle, to_onehot, flatten_arrays, from_flat, enzip  # noqa: F401

"""Observation / action spaces per version (reference utils/utils_env.py:9-68).  The
arithmetic of the version tables (ratios, rewards, action transforms) runs inside the
CUDA kernel; only the space definitions live on the host."""
import numpy as np

from custom_envs_b200.compat import spaces
from custom_envs_b200.utils.utils_common import History

_HISTORY_LAYOUTS = {           # version -> (uses max_history, keys in insertion order)
    0: (False, ('gradients',)),
    1: (True, ('losses', 'gradients')),
    2: (False, ('weights', 'losses', 'gradients')),
    3: (True, ('weights', 'losses', 'gradients')),
    4: (True, ('gradients',)),
    5: (True, ('weights', 'losses', 'gradients', 'actions')),
}


def history_layout(version, max_history):
    if version not in _HISTORY_LAYOUTS:
        raise RuntimeError()
    deep, keys = _HISTORY_LAYOUTS[version]
    return (max_history if deep else 1), keys


def get_obs_version(shape, max_history, version=0):
    depth, keys = history_layout(version, max_history)
    named = {key: (() if key == 'losses' else shape) for key in keys}
    space = spaces.Box(low=-1e6, high=1e6, dtype=np.float32, shape=(depth * len(keys),))
    return space, History(depth, **named)


def get_action_space_optlrs(version=0):
    bounds = {0: (-4., 6.), 1: (0., 1e4), 2: (-1e3, 1e4)}
    if version not in bounds:
        raise RuntimeError()
    low, high = bounds[version]
    return spaces.Box(low=low, high=high, dtype=np.float32, shape=(1,))

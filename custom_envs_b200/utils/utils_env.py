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


# ---------------------------------------------------------------------------------------------
# Host-side version tables (reference utils/utils_env.py:71-164).  The batched step evaluates the
# same formulas inside the CUDA kernels (csrc/b200env_shared.cuh: adjust_w/g/l, action_to_lr,
# step_scalars); these numpy forms exist for callers that import them directly and as the
# cross-check of the kernel's tables in tests/test_host_layer.py.
def _ratio_soft(new, old):
    return new / (np.abs(old) + 1e-3)


def _ratio_bare(new, old):
    with np.errstate(all='ignore'):
        return np.nan_to_num(new / np.abs(old))


_REWARDS = {
    0: lambda loss, adj: -float(adj),
    1: lambda loss, adj: float(1 / loss),
    2: lambda loss, adj: -float(adj) * 100,
    3: lambda loss, adj: float(1 / loss) * 100,
    4: lambda loss, adj: np.log(1 / loss),
    5: lambda loss, adj: -(float(adj) - 1) ** 2,
    6: lambda loss, adj: -(float(adj) - 1),
}

_LEARNING_RATES = {
    0: lambda action: 10 ** (action - 4),
    1: lambda action: action * 1e-3,
    2: lambda action: 2 ** action,
    3: lambda action: np.clip((action + 1e3) * 1e-6, 0, np.inf),
}


def get_reward(loss, adjusted_loss, version=0):
    """Reward of one step from the loss / adjusted loss; RuntimeError on an unknown version."""
    if version not in _REWARDS:
        raise RuntimeError()
    return _REWARDS[version](loss, adjusted_loss)


def get_action_optlrs(action, version):
    """Agent action -> learning rate for the OptLRs family; RuntimeError on an unknown version."""
    if version not in _LEARNING_RATES:
        raise RuntimeError()
    return _LEARNING_RATES[version](action)


def get_observation(history, version=0):
    """(adjusted loss: float, adjusted weights [P], adjusted gradients [P]) from the raw History
    (entries newest first).  Versions 0 and 1 damp the denominators with 1e-3, version 1 scales the
    raw gradient instead of a ratio, version 2 compares consecutive differences, version 3 is the
    bare ratio made finite with nan_to_num (the one MultiOptLRs uses)."""
    losses, grads, weights = history['losses'], history['gradients'], history['weights']
    if version in (0, 1):
        adj_loss, adj_wght = _ratio_soft(losses[0], losses[1]), _ratio_soft(weights[0], weights[1])
        adj_grad = _ratio_soft(grads[0], grads[1]) if version == 0 else grads[0] * 1e2
    elif version == 2:
        adj_loss = (losses[0] - losses[1]) / (np.abs(losses[1] - losses[2]) + 1e-3)
        adj_wght = np.abs(weights[1] - weights[2]) / (np.abs(weights[0] - weights[1]) + 1e-8)
        adj_grad = (grads[0] - grads[1]) / (np.abs(grads[1] - grads[2]) + 1e-3)
    elif version == 3:
        adj_loss, adj_wght = _ratio_bare(losses[0], losses[1]), _ratio_bare(weights[0], weights[1])
        adj_grad = _ratio_bare(grads[0], grads[1])
    else:
        raise RuntimeError()
    return float(adj_loss), adj_wght, adj_grad

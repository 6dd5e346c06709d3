"""MultiOptimize: one agent per parameter chooses that parameter's update
(reference envs/multioptimize.py:19-163); the step runs in libb200env.so.

The reference constructor cannot build its problem at HEAD (``get_problem(data_set=...)``
falls through to the Rosenbrock problem, which rejects ``data_set``; envs/multioptimize.py:44,
problems/__init__.py:7-16).  The intended problem -- the 'nn' classifier on
``load_data(data_set, batch_size)`` -- is what is built here."""
from custom_envs_b200.compat import spaces
from custom_envs_b200.envs.device_env import DeviceEnvFront, VersionType
from custom_envs_b200.problems import get_problem
from custom_envs_b200.utils import utils_env

import numpy as np


class MultiOptimize(DeviceEnvFront):
    """``MultiOptimize(data_set='iris', batch_size=None, version=1, max_batches=400,
    max_history=5, observation_version=0, action_version=0, reward_version=0)`` as in the
    reference; ``layers`` (hidden widths of the classifier) and ``device`` are extensions."""
    ENV_KIND = 'optimize'

    def __init__(self, data_set='iris', batch_size=None, version=1, max_batches=400,
                 max_history=5, observation_version=0, action_version=0, reward_version=0,
                 layers=None, device='cuda:0'):
        super().__init__()
        if hasattr(data_set, 'spec'):
            model = data_set
        else:
            from custom_envs_b200.data import load_data
            data = load_data(data_set, batch_size) if isinstance(data_set, str) else data_set
            model = get_problem('nn', data_set=data, layers=layers)
        model.device = device
        obs_space, _ = utils_env.get_obs_version((model.size,), max_history, version)
        if version == 5:            # History.append asserts on the missing 'weights' key
            raise RuntimeError('history version 5 cannot be fed by MultiOptimize '
                               '(AssertionError in the reference, utils/utils_common.py:183)')
        if action_version == 0:                                    # multioptimize.py:51-58
            low, high = -4., 4.
        elif action_version == 1:
            low, high = -1e8, 1e8
        else:
            raise RuntimeError()
        if observation_version not in (0, 1, 2, 3) or reward_version not in range(7):
            raise RuntimeError()
        act_space = spaces.Box(low=low, high=high, dtype=np.float32, shape=(1,))
        self._setup(model, obs_space, act_space, max_batches, max_history,
                    VersionType(version, observation_version, action_version, reward_version),
                    device)

    def __repr__(self):
        return '<MultiOptimize({})>'.format(self.version)

    def _terminal(self):
        return self.current_step >= self.max_batches

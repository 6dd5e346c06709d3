"""Observation wrappers for dict-of-agents envs (reference wrappers/optimizewrappers.py:9-70)."""
import numpy as np

from custom_envs_b200.compat import Wrapper, spaces
from custom_envs_b200.utils.utils_common import History


class HistoryWrapper(Wrapper):
    """Every agent observes its last ``max_history`` observations (newest first)."""

    def __init__(self, env, max_history=5):
        named_shapes = {key: space.shape for key, space in env.observation_space.spaces.items()}
        env.observation_space = spaces.Dict({
            key: spaces.Box(low=np.array([space.low] * max_history),
                            high=np.array([space.high] * max_history), dtype=space.dtype)
            for key, space in env.observation_space.spaces.items()})
        self.history = History(max_history, **named_shapes)
        super().__init__(env)

    def step(self, action):
        state, reward, terminal, info = self.env.step(action)
        self.history.append(**state)
        return dict(self.history), reward, terminal, info

    def reset(self, **kwargs):
        self.history.reset(**self.env.reset(**kwargs))
        return dict(self.history)

    def __repr__(self):
        return '<{}{!r}{!r}>'.format(type(self).__name__, self.history, self.env)


class SubSetWrapper(Wrapper):
    """Keep only the agents named in ``subset``."""

    def __init__(self, env, subset):
        env.observation_space = spaces.Dict({key: env.observation_space[key] for key in subset})
        self.subset = subset
        super().__init__(env)

    def step(self, action):
        state, reward, terminal, info = self.env.step(action)
        return {name: state[name] for name in self.subset}, reward, terminal, info

    def reset(self, **kwargs):
        state = self.env.reset(**kwargs)
        return {name: state[name] for name in self.subset}

    def __repr__(self):
        return '<{}{!r}{!r}>'.format(type(self).__name__, self.subset, self.env)

"""Environment base classes (reference envs/baseenvironment.py:11-64)."""
from custom_envs_b200.compat import Env, np_random
from custom_envs_b200.utils.utils_math import use_random_state


class BaseEnvironment(Env):
    def __init__(self):
        self.random_generator, _ = np_random()
        self.current_step = 0

    def seed(self, seed=None):
        self.random_generator, _ = np_random(seed)

    def step(self, action):
        self.current_step += 1
        with use_random_state(self.random_generator):
            state, reward, terminal, info = self.base_step(action)
        info['episode'] = {'r': reward, 'l': self.current_step}
        return state, reward, terminal, info

    def reset(self):
        self.current_step = 0
        with use_random_state(self.random_generator):
            return self.base_reset()

    def base_step(self, action):
        raise NotImplementedError

    def base_reset(self):
        raise NotImplementedError


class BaseMultiEnvironment(BaseEnvironment):
    AGENT_FMT = 'parameter-{:d}'

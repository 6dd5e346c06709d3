"""Base classes of the gym-facing envs.

Contract kept from reference envs/baseenvironment.py:11-64: subclasses implement
``base_reset() -> state`` and ``base_step(action) -> (state, reward, terminal, info)``; the public
``reset`` / ``step`` run them with numpy's GLOBAL random state swapped for the env's own generator
(so that ``seed()`` makes data shuffles reproducible), count steps in ``current_step`` and attach
``info['episode'] = {'r': reward of this step, 'l': steps so far}``."""
from custom_envs_b200.compat import Env, np_random
from custom_envs_b200.utils.utils_math import use_random_state


class BaseEnvironment(Env):
    def __init__(self):
        self.current_step = 0
        self.seed()

    def seed(self, seed=None):
        self.random_generator = np_random(seed)[0]

    def _with_own_rng(self, call, *args):
        with use_random_state(self.random_generator):
            return call(*args)

    def reset(self):
        self.current_step = 0
        return self._with_own_rng(self.base_reset)

    def step(self, action):
        self.current_step += 1
        outcome = tuple(self._with_own_rng(self.base_step, action))
        outcome[3]['episode'] = dict(r=outcome[1], l=self.current_step)
        return outcome

    def base_reset(self):
        raise NotImplementedError('%s.base_reset' % type(self).__name__)

    def base_step(self, action):
        raise NotImplementedError('%s.base_step' % type(self).__name__)


class BaseMultiEnvironment(BaseEnvironment):
    """Envs whose observation / action are dicts keyed by agent name."""
    AGENT_FMT = 'parameter-{:d}'

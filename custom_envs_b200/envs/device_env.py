"""Gym front of a device-resident optimise env.

Stand-alone, a front owns a one-env ``BatchedOptEnv`` (natural row order) and behaves like
the reference env object: dict-of-agents observations, float reward, bool done, info dict.
Inside ``OptVecEnv`` all fronts of one configuration are fused into ONE backend
(``fuse_fronts``); the front then only mirrors ``current_step`` and replays its env's
per-step result to wrappers (``Monitor``)."""
from collections import OrderedDict, namedtuple

import numpy as np

from custom_envs_b200.batched_env import BatchedOptEnv, env_permutations_device
from custom_envs_b200.compat import spaces
from custom_envs_b200.envs.baseenvironment import BaseMultiEnvironment

VersionType = namedtuple('VersionType', ['history', 'observation', 'action', 'reward'])


class LazyAgentDict(dict):
    """Marker type: the dict-of-agents view of rows that already sit in the VecEnv buffer."""


class DeviceEnvFront(BaseMultiEnvironment):
    ENV_KIND = 'optlrs'

    def _setup(self, model, obs_space, act_space, max_batches, max_history, version, device):
        self.model = model
        self.max_history = max_history
        self.max_batches = max_batches
        self.version = version
        self.device = device
        names = [self.AGENT_FMT.format(i) for i in range(model.size)]
        self._names = names
        self.observation_space = spaces.Dict({name: obs_space for name in names})
        self.action_space = spaces.Dict({name: act_space for name in names})
        self._backend = None
        self._slot = 0
        self._attached = False
        self._pending = None
        self.seed()

    # ------------------------------------------------------------ backend plumbing
    def backend_kwargs(self):
        return dict(batch_size=self.model.batch_size, max_batches=self.max_batches,
                    max_history=self.max_history, env_kind=self.ENV_KIND,
                    history_version=self.version.history, observation_version=self.version.observation,
                    action_version=self.version.action, reward_version=self.version.reward)

    def fuse_key(self):
        """Envs with equal keys can live in one ``BatchedOptEnv``: same env class, problem shape,
        versions and the same data-set CONTENT (a digest, computed once per data-set object)."""
        return (type(self).__name__, self.model.spec, self.model.data_key(),
                tuple(sorted(self.backend_kwargs().items())))

    def seed(self, seed=None):
        super().seed(seed)
        if getattr(self, '_backend', None) is not None and not self._attached:
            self._backend.close()             # permutation depends on the seed: rebuild lazily
            self._backend = None

    def _attach(self, backend, slot):
        if self._backend is not None and not self._attached:
            self._backend.close()
        self._backend, self._slot, self._attached = backend, slot, True
        self.model._bind(backend, slot)

    def _standalone(self):
        if self._backend is None:
            feats, targs = self.model.device_arrays()
            perms = None if feats is None else env_permutations_device(len(feats), [self.random_generator], self.device)
            self._backend = BatchedOptEnv(self.model.spec, feats, targs, 1, row_order='natural',
                                          auto_reset=False, perms=perms, device=self.device,
                                          **self.backend_kwargs())
            self.model._bind(self._backend, 0)
        return self._backend

    # inside a fused OptVecEnv the step counters of all envs live in one numpy array that the
    # VecEnv updates once per step; the attribute of the reference (baseenvironment.py:13,37)
    # reads through to it
    _shared_steps = None

    @property
    def current_step(self):
        if self._shared_steps is not None:
            steps, slot = self._shared_steps
            return int(steps[slot])
        return self._current_step

    @current_step.setter
    def current_step(self, value):
        self._current_step = value
        if self._shared_steps is not None:
            steps, slot = self._shared_steps
            steps[slot] = value

    def _host_reset_done(self):
        self.current_step = 0

    def _host_step_done(self, info_row, done):
        self.current_step = 0 if done else int(info_row[15])

    def _agent_dict(self, rows):
        return OrderedDict((name, rows[i]) for i, name in enumerate(self._names))

    # ----------------------------------------------------------------- gym surface
    def reset(self):
        if self._attached:
            self.current_step = 0
            return LazyAgentDict()
        self.current_step = 0
        obs = self._standalone().reset().cpu().numpy()
        return self._agent_dict(obs)

    def step(self, action):
        from custom_envs_b200.vectorize.optvecenv import info_row_to_dict
        if self._attached:
            reward, done, info = self._pending
            self._pending = None
            return LazyAgentDict(), reward, done, info
        import torch
        backend = self._standalone()
        flat = np.reshape([np.ravel(action[name]) for name in self._names], (-1,)).astype(np.float32)
        obs, reward, done, info = backend.step(torch.as_tensor(flat, device=backend.device))
        info_row = info[0].cpu().numpy()
        self.current_step = int(info_row[15])
        # the reward of the step is the double in info (info['episode']['r'] is the very same
        # object in the reference, baseenvironment.py:40); reward_out is its float32 copy
        return (self._agent_dict(obs.cpu().numpy()), float(info_row[14]), bool(done[0].item()),
                info_row_to_dict(info_row))

    def render(self, mode='human'):
        pass

    def close(self):
        if self._backend is not None and not self._attached:
            self._backend.close()
            self._backend = None


def fuse_fronts(fronts, device=None):
    """One ``BatchedOptEnv`` for all fronts (VecEnv row order, auto-reset on)."""
    first = fronts[0]
    feats, targs = first.model.device_arrays()
    perms = None
    if feats is not None:
        perms = env_permutations_device(len(feats), [front.random_generator for front in fronts], device or first.device)
    backend = BatchedOptEnv(first.model.spec, feats, targs, len(fronts), row_order='lexicographic',
                            auto_reset=True, perms=perms, device=device or first.device,
                            **first.backend_kwargs())
    for slot, front in enumerate(fronts):
        front._attach(backend, slot)
    return backend

"""Observation wrappers for the dict-of-agents envs, host and device side.

``HistoryWrapper`` / ``SubSetWrapper`` keep the constructor arguments and the observable
behaviour of reference wrappers/optimizewrappers.py:9-70 (gym ``Wrapper`` around ONE env, agent
dict in, agent dict out).  ``DeviceHistoryWrapper`` / ``DeviceSubSetWrapper`` do the same to the
``[E*P, obs_dim]`` observation matrix of a ``BatchedOptEnv`` without leaving HBM,
so that wrapping a fused env batch costs no host round trip."""
import numpy as np

from custom_envs_b200.compat import Wrapper, spaces


class _ObservationRewriter(Wrapper):
    """``step`` and ``reset`` funnel the agent dict through one hook, ``rewrite(state, fresh)``;
    ``fresh`` is true for the first observation of an episode."""

    def rewrite(self, state, fresh):
        raise NotImplementedError

    def reset(self, **kwargs):
        return self.rewrite(self.env.reset(**kwargs), True)

    def step(self, action):
        outcome = self.env.step(action)
        return (self.rewrite(outcome[0], False),) + tuple(outcome[1:])


class HistoryWrapper(_ObservationRewriter):
    """Each agent sees its ``max_history`` latest observations stacked on a new leading axis,
    newest at index 0.  A reset fills every slot with the first observation of the episode."""

    def __init__(self, env, max_history=5):
        depth = int(max_history)
        inner = env.observation_space.spaces
        self.max_history = depth
        self._stack = {name: np.zeros((depth,) + tuple(box.shape) if box.shape else (depth, 1))
                       for name, box in inner.items()}
        env.observation_space = spaces.Dict({
            name: spaces.Box(low=np.stack([box.low] * depth), high=np.stack([box.high] * depth), dtype=box.dtype)
            for name, box in inner.items()})
        super().__init__(env)

    def rewrite(self, state, fresh):
        assert state.keys() == self._stack.keys()
        for name, rows in self._stack.items():
            newest = np.reshape(state[name], rows.shape[1:])
            if fresh:
                rows[:] = newest
            else:
                rows[1:] = rows[:-1].copy()
                rows[0] = newest
        return {name: rows.copy() for name, rows in self._stack.items()}

    def __repr__(self):
        shapes = {name: rows.shape[1:] for name, rows in self._stack.items()}
        return '<{}<max_history={}, shapes={!r}>{!r}>'.format(type(self).__name__, self.max_history, shapes, self.env)


class SubSetWrapper(_ObservationRewriter):
    """Only the agents listed in ``subset`` stay in the observation (and its space)."""

    def __init__(self, env, subset):
        self.subset = subset
        env.observation_space = spaces.Dict({name: env.observation_space[name] for name in subset})
        super().__init__(env)

    def rewrite(self, state, fresh):
        return {name: state[name] for name in self.subset}

    def __repr__(self):
        return '<{}{!r}{!r}>'.format(type(self).__name__, self.subset, self.env)


# ----------------------------------------------------------------------------- device side
class _DeviceRewriter:
    """Same hook over a ``BatchedOptEnv``-like object: ``reset() -> obs`` and ``step(actions) ->
    (obs, reward, done, info)`` with ``obs`` a ``[E*P, dim]`` device tensor whose rows of env ``e``
    are ``[e*P, (e+1)*P)``.  Envs that finish are reset inside ``step`` by the fused env (their rows
    already hold the first observation of the next episode), which is when the reference's worker
    calls ``wrapper.reset()``: ``rewrite`` gets those envs as ``fresh_envs`` (uint8 ``[E]``)."""

    def __init__(self, env):
        self.env = env
        self.num_envs, self.num_params = env.num_envs, env.num_params
        self.device = env.device

    def __getattr__(self, name):
        return getattr(self.env, name)

    def rewrite(self, obs, fresh_envs):
        raise NotImplementedError

    def reset(self, *args, **kwargs):
        return self.rewrite(self.env.reset(*args, **kwargs), None)

    def step(self, actions, *args, **kwargs):
        obs, reward, done, info = self.env.step(actions, *args, **kwargs)
        return self.rewrite(obs, done), reward, done, info


class DeviceHistoryWrapper(_DeviceRewriter):
    """``HistoryWrapper`` for a whole env batch: a ring of the ``max_history`` latest observation
    matrices in HBM; returns ``[E*P, max_history, dim]``, newest first."""

    def __init__(self, env, max_history=5):
        import torch
        super().__init__(env)
        self._torch = torch
        self.max_history = int(max_history)
        self.obs_dim = env.obs_dim
        self._ring = torch.zeros((self.max_history, env.num_rows, env.obs_dim), dtype=torch.float32,
                                 device=env.device)
        self._head = 0                      # slot of the newest matrix
        self._slots = torch.arange(self.max_history, device=env.device)

    def rewrite(self, obs, fresh_envs):
        torch = self._torch
        if fresh_envs is None:
            self._ring[:] = obs                                   # every slot = first observation
        else:
            self._head = (self._head - 1) % self.max_history
            self._ring[self._head] = obs
            rows = fresh_envs.to(torch.bool).repeat_interleave(self.num_params)
            if bool(rows.any()):
                self._ring[:, rows] = obs[rows]
        order = (self._slots + self._head) % self.max_history     # newest first
        return self._ring.index_select(0, order).transpose(0, 1)

    def __repr__(self):
        return '<{}<max_history={}>{!r}>'.format(type(self).__name__, self.max_history, self.env)


class DeviceSubSetWrapper(_DeviceRewriter):
    """``SubSetWrapper`` for a whole env batch.  ``subset`` names agents (``'parameter-12'``) or
    gives parameter numbers; the result keeps, env by env, the rows of those agents in the order
    of ``subset``: ``[E*len(subset), dim]``."""

    def __init__(self, env, subset, row_order='lexicographic'):
        import torch
        super().__init__(env)
        params = [int(str(item).rsplit('-', 1)[-1]) for item in subset]
        assert all(0 <= p < env.num_params for p in params)
        if row_order == 'lexicographic':    # rows of an env are its agent names in string order
            names = sorted('parameter-%d' % p for p in range(env.num_params))
            row_of = {int(name.rsplit('-', 1)[-1]): row for row, name in enumerate(names)}
        else:
            row_of = {p: p for p in params}
        local = torch.tensor([row_of[p] for p in params], dtype=torch.int64, device=env.device)
        base = torch.arange(env.num_envs, dtype=torch.int64, device=env.device) * env.num_params
        self.subset = list(subset)
        self.rows = (base[:, None] + local[None, :]).reshape(-1)

    def rewrite(self, obs, fresh_envs):
        return obs.index_select(0, self.rows)

    def __repr__(self):
        return '<{}{!r}{!r}>'.format(type(self).__name__, self.subset, self.env)

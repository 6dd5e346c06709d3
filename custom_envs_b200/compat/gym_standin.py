"""Minimal stand-ins for the slice of classic ``gym`` the reference boundary touches.

``gym`` is not installable in the build image (no network), so the package falls back
to these when ``import gym`` fails (see ``compat/__init__.py``).  Only behaviour the
reference relies on is provided: ``spaces.Box`` / ``spaces.Dict`` (sorted keys, as
classic gym does for plain dicts -- this is what makes the agent rows lexicographic,
reference envs/multioptlrs.py:50-57), ``Env``, ``Wrapper``, ``utils.seeding.np_random``
(classic gym returned a ``numpy.random.RandomState``; reference
envs/baseenvironment.py:17,28 depends on that) and ``register`` / ``make``.
"""
from __future__ import annotations

import hashlib
import importlib
import os
import struct
import sys
import types
from collections import OrderedDict

import numpy as np


# ------------------------------------------------------------------------- seeding
def _bigint_from_bytes(data: bytes) -> int:
    pad = (4 - len(data) % 4)
    data = data + b'\0' * pad
    words = struct.unpack('%dI' % (len(data) // 4), data)
    return sum(word << (32 * i) for i, word in enumerate(words))


def create_seed(seed=None, max_bytes=8):
    if seed is None:
        return _bigint_from_bytes(os.urandom(max_bytes))
    if isinstance(seed, str):
        digest = hashlib.sha512(seed.encode('utf8')).digest()
        return _bigint_from_bytes(digest[:max_bytes])
    if isinstance(seed, (int, np.integer)):
        return int(seed) % 2 ** (8 * max_bytes)
    raise TypeError('Invalid type for seed: %r' % type(seed))


def hash_seed(seed=None, max_bytes=8):
    if seed is None:
        seed = create_seed(max_bytes=max_bytes)
    digest = hashlib.sha512(str(seed).encode('utf8')).digest()
    return _bigint_from_bytes(digest[:max_bytes])


def np_random(seed=None):
    """Classic-gym seeding: a RandomState seeded with the 32-bit limbs of a hashed seed."""
    if seed is not None and not (isinstance(seed, (int, np.integer)) and seed >= 0):
        raise ValueError('Seed must be a non-negative integer or omitted, not %r' % seed)
    seed = create_seed(seed)
    limbs, big = [], hash_seed(seed)
    while big > 0:
        big, low = divmod(big, 2 ** 32)
        limbs.append(low)
    rng = np.random.RandomState()
    rng.seed(limbs or [0])
    return rng, seed


# -------------------------------------------------------------------------- spaces
class Space:
    def __init__(self, shape=None, dtype=None):
        self.shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)
        self.np_random, _ = np_random()

    def seed(self, seed=None):
        self.np_random, seed = np_random(seed)
        return [seed]

    def __contains__(self, item):
        return self.contains(item)


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        dtype = np.dtype(dtype)
        if shape is None:
            low, high = np.asarray(low), np.asarray(high)
            shape = low.shape
        else:
            shape = tuple(int(s) for s in shape)
            low = np.full(shape, low) if np.isscalar(low) else np.asarray(low)
            high = np.full(shape, high) if np.isscalar(high) else np.asarray(high)
        self.low = low.astype(dtype)
        self.high = high.astype(dtype)
        super().__init__(shape, dtype)

    def sample(self):
        low = self.low.astype(np.float64)
        high = self.high.astype(np.float64)
        return self.np_random.uniform(low=low, high=high, size=self.shape).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return (x.shape == self.shape and bool(np.all(x >= self.low))
                and bool(np.all(x <= self.high)))

    def __eq__(self, other):
        return (isinstance(other, Box) and self.shape == other.shape
                and np.allclose(self.low, other.low) and np.allclose(self.high, other.high))

    def __repr__(self):
        return 'Box%s' % (self.shape,)


class Dict(Space):
    def __init__(self, spaces=None, **spaces_kwargs):
        if spaces is None:
            spaces = spaces_kwargs
        if isinstance(spaces, dict) and not isinstance(spaces, OrderedDict):
            spaces = OrderedDict(sorted(list(spaces.items())))
        elif isinstance(spaces, (list, tuple)):
            spaces = OrderedDict(spaces)
        self.spaces = spaces
        super().__init__(None, None)

    def seed(self, seed=None):
        return [space.seed(seed) for space in self.spaces.values()]

    def sample(self):
        return OrderedDict((key, space.sample()) for key, space in self.spaces.items())

    def contains(self, x):
        if not isinstance(x, dict) or len(x) != len(self.spaces):
            return False
        return all(key in x and space.contains(x[key])
                   for key, space in self.spaces.items())

    def __getitem__(self, key):
        return self.spaces[key]

    def __iter__(self):
        return iter(self.spaces)

    def __len__(self):
        return len(self.spaces)

    def __eq__(self, other):
        return isinstance(other, Dict) and self.spaces == other.spaces

    def __repr__(self):
        return 'Dict(%d spaces)' % len(self.spaces)


class Tuple(Space):
    def __init__(self, spaces):
        self.spaces = tuple(spaces)
        super().__init__(None, None)

    def sample(self):
        return tuple(space.sample() for space in self.spaces)

    def contains(self, x):
        return (isinstance(x, (tuple, list)) and len(x) == len(self.spaces)
                and all(s.contains(p) for s, p in zip(self.spaces, x)))


# ----------------------------------------------------------------------------- core
class Env:
    metadata = {'render.modes': []}
    reward_range = (-float('inf'), float('inf'))
    spec = None
    action_space = None
    observation_space = None

    def step(self, action):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    def render(self, mode='human'):
        raise NotImplementedError

    def close(self):
        pass

    def seed(self, seed=None):
        return

    @property
    def unwrapped(self):
        return self

    def __enter__(self):
        return self

    def __exit__(self, *args):
        self.close()
        return False


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self.action_space = self.env.action_space
        self.observation_space = self.env.observation_space
        self.reward_range = self.env.reward_range
        self.metadata = self.env.metadata

    def __getattr__(self, name):
        if name.startswith('_'):
            raise AttributeError("attempted to get missing private attribute '%s'" % name)
        return getattr(self.env, name)

    def step(self, action):
        return self.env.step(action)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def render(self, mode='human', **kwargs):
        return self.env.render(mode, **kwargs)

    def close(self):
        return self.env.close()

    def seed(self, seed=None):
        return self.env.seed(seed)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def __repr__(self):
        return '<%s%r>' % (type(self).__name__, self.env)


# --------------------------------------------------------------------- registration
_REGISTRY = {}


def register(id, entry_point=None, **kwargs):   # noqa: A002 (gym's own argument name)
    _REGISTRY[id] = (entry_point, kwargs.get('kwargs', {}))


def make(id, **kwargs):                         # noqa: A002
    if id not in _REGISTRY:
        raise KeyError('No registered env with id: %s' % id)
    entry_point, defaults = _REGISTRY[id]
    if isinstance(entry_point, str):
        mod_name, attr = entry_point.split(':')
        entry_point = getattr(importlib.import_module(mod_name), attr)
    merged = dict(defaults)
    merged.update(kwargs)
    return entry_point(**merged)


def install_as_gym():
    """Expose these stand-ins under the module names the reference imports."""
    this = sys.modules[__name__]
    gym = types.ModuleType('gym')
    gym.Env, gym.Wrapper, gym.make, gym.register = Env, Wrapper, make, register
    spaces = types.ModuleType('gym.spaces')
    spaces.Space, spaces.Box, spaces.Dict, spaces.Tuple = Space, Box, Dict, Tuple
    core = types.ModuleType('gym.core')
    core.Env, core.Wrapper = Env, Wrapper
    utils = types.ModuleType('gym.utils')
    seeding = types.ModuleType('gym.utils.seeding')
    seeding.np_random, seeding.hash_seed, seeding.create_seed = np_random, hash_seed, create_seed
    utils.seeding = seeding
    envs = types.ModuleType('gym.envs')
    registration = types.ModuleType('gym.envs.registration')
    registration.register, registration.make = register, make
    envs.registration = registration
    gym.spaces, gym.core, gym.utils, gym.envs = spaces, core, utils, envs
    gym.__standin__ = this
    for name, mod in (('gym', gym), ('gym.spaces', spaces), ('gym.core', core),
                      ('gym.utils', utils), ('gym.utils.seeding', seeding),
                      ('gym.envs', envs), ('gym.envs.registration', registration)):
        sys.modules.setdefault(name, mod)
    return sys.modules['gym']

"""CPU tests of the host-side mirror of the reference interfaces (no GPU, no compute calls).
They re-express the reference's own tests: tests/utils/test_utils_common.py,
tests/utils/test_utils_env.py, tests/utils/test_utils_math.py,
tests/dataset/test_inmemorydataset.py, tests/vectorize/*, tests/wrappers/*."""
import ctypes
import os
import tempfile
from functools import partial
from itertools import chain

import numpy as np
import numpy.random as npr
import pandas as pd
import pytest

import custom_envs                                     # noqa: F401  (alias package)
from custom_envs.utils import utils_common, utils_env, utils_math
from custom_envs.dataset import InMemoryDataSet
from custom_envs.vectorize import OptVecEnv, SubprocVecEnv, ThreadVecEnv
from custom_envs.vectorize.optvecenv import flatten_dictionary
from custom_envs.wrappers import HistoryWrapper, SubSetWrapper
from custom_envs.wrappers.monitor import Monitor as StrictMonitor
from custom_envs.utils.utils_logging import Monitor
from custom_envs_b200.compat import gym, spaces
from custom_envs_b200 import _lib

Box, Dict = spaces.Box, spaces.Dict


class StubEnv(gym.Env):
    """Dict-of-agents stub with a 10 step episode (as in the reference's tests)."""

    def __init__(self):
        self.counter = 0
        self.observation_space = Dict({name: Box(low=-1e3, high=1e3, dtype=np.float32, shape=[5])
                                       for name in ('test1d', 'test2d', 'test3d')})
        self.action_space = self.observation_space
        self._sample_space = self.observation_space       # wrappers replace observation_space

    def step(self, action):
        self.counter += 1
        return self._sample_space.sample(), 0, self.counter >= 10, {}

    def reset(self):
        self.counter = 0
        return self._sample_space.sample()

    def render(self, mode='human'):
        pass


class ArrayStubEnv(StubEnv):
    def __init__(self):
        super().__init__()
        self.observation_space = Dict({
            'test1d': Box(low=-1e3, high=1e3, dtype=np.float32, shape=[5]),
            'test2d': Box(low=-1e3, high=1e3, dtype=np.float32, shape=[5] * 2),
            'test3d': Box(low=-1e3, high=1e3, dtype=np.float32, shape=[5] * 3)})
        self._sample_space = self.observation_space
        self.action_space = Box(low=-1e3, high=1e3, dtype=np.float32, shape=(25,))


# ------------------------------------------------------------------ C ABI presence
def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    include = os.path.join(os.path.dirname(__file__), '..', 'include')
    header = ''.join(open(os.path.join(include, name)).read() for name in sorted(os.listdir(include)))
    import re
    declared = set(re.findall(r'\b(b2[edp]_[a-z_]+)\s*\(', header))
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.b2e_abi_version() == 2
    assert ctypes.sizeof(_lib.Config) == 120


def test_create_fails_loudly_without_gpu_or_on_bad_config():
    lib = _lib.load()
    cfg = _lib.Config()
    handle = ctypes.c_void_p()
    assert lib.b2e_create(ctypes.byref(cfg), ctypes.byref(handle)) != 0      # struct_size == 0
    assert b'struct_size' in lib.b2e_last_error(None)


# ------------------------------------------------------------------------ History
def test_history_build_multistate_is_ravel():
    test_a, test_b = npr.rand(5, 5), npr.rand(1)
    history = utils_common.History(1, test_a=(5, 5), test_b=(1,))
    history.append(test_a=test_a, test_b=test_b)
    hist_a, hist_b = zip(*history.build_multistate())
    assert np.all(hist_a == test_a.ravel()) and np.all(hist_b == test_b.ravel())


def test_history_newest_first_and_reset():
    depth = 3
    history = utils_common.History(depth, test1d=(5,), test2d=(5, 5))
    assert history['test1d'].shape == (depth, 5) and history['test2d'].shape == (depth, 5, 5)
    test1d = np.arange(depth * 5).reshape((depth, 5))
    test2d = np.arange(depth * 25).reshape((depth, 5, 5))
    for i in range(depth):
        history.append(test1d=test1d[-(i + 1)], test2d=test2d[-(i + 1)])
    assert np.all(history['test1d'] == test1d) and np.all(history['test2d'] == test2d)
    history.reset()
    assert np.all(history['test1d'] == 0)
    history.reset(test1d=test1d[0], test2d=test2d[0])
    assert np.all(history['test1d'] == test1d[0]) and np.all(history['test2d'] == test2d[0])
    assert list(history) == ['test1d', 'test2d'] and len(history) == 2


def test_history_multistate_shapes():
    history = utils_common.History(4, weights=(6,), losses=(), gradients=(6,))
    states = history.build_multistate()
    assert len(states) == 6 and all(len(s) == 12 for s in states)


def test_flatten_roundtrip_and_shuffle_alignment():
    arrays = [npr.rand(3, 4), npr.rand(4), npr.rand(2, 2, 2)]
    flat = utils_common.flatten_arrays(arrays)
    assert flat.dtype == np.float64 and flat.shape == (24,)
    for a, b in zip(arrays, utils_common.from_flat(flat, [a.shape for a in arrays])):
        assert np.array_equal(a, b)
    a, b = utils_common.shuffle(np.arange(25, 0, -1), np.arange(25, 0, -1))
    assert np.all(a == b)
    onehot, n = utils_common.to_onehot(np.arange(30) % 7)
    assert n == 7 and onehot.shape == (30, 7) and len(np.unique(onehot, axis=0)) == 7


# --------------------------------------------------------------- utils_env / math
@pytest.mark.parametrize('version,dim', [(0, 1), (1, 10), (2, 3), (3, 15), (4, 5)])
def test_obs_space_shapes(version, dim):
    space, history = utils_env.get_obs_version((7,), 5, version)
    assert space.shape == (dim,) and space.dtype == np.float32
    assert np.all(space.low == -1e6) and np.all(space.high == 1e6)
    assert len(history.build_multistate()) == 7


def test_bad_versions_raise_runtime_error():
    with pytest.raises(RuntimeError):
        utils_env.get_obs_version((3,), 5, 9)
    with pytest.raises(RuntimeError):
        utils_env.get_action_space_optlrs(7)


def test_action_space_bounds():
    assert (utils_env.get_action_space_optlrs(0).low, utils_env.get_action_space_optlrs(0).high) == (-4., 6.)
    assert (utils_env.get_action_space_optlrs(1).low, utils_env.get_action_space_optlrs(1).high) == (0., 1e4)
    assert (utils_env.get_action_space_optlrs(2).low, utils_env.get_action_space_optlrs(2).high) == (-1e3, 1e4)


@pytest.mark.parametrize('seed', range(3))
def test_use_random_state(seed):
    state = npr.RandomState(seed)
    with utils_math.use_random_state(state):
        inside = tuple(npr.rand() for _ in range(20))
    fresh = npr.RandomState(seed)
    assert inside == tuple(fresh.rand() for _ in range(20))
    with utils_math.use_random_state(state):           # the env's generator never advances
        assert tuple(npr.rand() for _ in range(20)) == inside
    data = utils_math.normalize(npr.randn(8, 4))
    assert data.min() >= 0 and data.max() <= 1


# ------------------------------------------------------------------------ dataset
def test_inmemorydataset_batches():
    feats, targs = np.arange(20).reshape(10, 2), np.arange(10)
    data = InMemoryDataSet(feats, targs, 4)
    assert len(data) == 3 and [len(b.features) for b in data] == [4, 4, 2]
    assert len(InMemoryDataSet(feats, targs, None)) == 1
    assert data.feature_shape == (2,) and data.target_shape == ()
    data.on_epoch_end()
    assert sorted(data.targets) == list(range(10))
    assert np.all(data.features[:, 0] // 2 == data.targets)


def test_load_data_iris_and_synthetic():
    from custom_envs import load_data
    iris = load_data('iris', 32)
    assert iris.features.shape == (150, 4) and iris.targets.shape == (150, 3)
    assert iris.features.min() >= 0 and iris.features.max() <= 1 and len(iris) == 5
    with pytest.warns(UserWarning):
        mnist = load_data('mnist-test', 32)
    assert mnist.features.shape == (10000, 49) and mnist.targets.shape == (10000, 10)
    with pytest.raises(RuntimeError):
        load_data('nope')


# ---------------------------------------------------------------------- vectorize
@pytest.mark.parametrize('vec_cls', [ThreadVecEnv, SubprocVecEnv])
def test_concurrent_vec_env(vec_cls):
    vec = vec_cls([partial(ArrayStubEnv) for _ in range(2)])
    states = vec.reset()
    assert len(states['test1d']) == 2
    done = False
    steps = 0
    while not done:
        states, rewards, dones, infos = vec.step([ArrayStubEnv().action_space.sample()] * 2)
        assert states['test3d'].shape == (2, 5, 5, 5) and len(rewards) == 2 and len(infos) == 2
        done = bool(np.any(dones))
        steps += 1
    assert steps == 10
    assert vec.get_attr('counter') == [0, 0]              # auto-reset happened in the worker
    vec.set_attr('counter', 4)
    assert vec.env_method('reset') is not None
    vec.close()
    vec.close()                                           # idempotent


def test_optvecenv_generic_path_rows():
    vec = OptVecEnv([StubEnv] * 2)
    assert not vec.is_device_backed and vec.num_envs == 6 and vec.agent_no_list == [3, 3]
    assert vec.reset().shape == (6, 5)
    done = False
    while not done:
        actions = list(chain.from_iterable([flatten_dictionary(StubEnv().action_space.sample())] * 2))
        states, rewards, terminals, infos = vec.step(actions)
        assert len(states) == vec.num_envs == len(rewards) == len(terminals) == len(infos)
        done = bool(np.any(terminals))
    vec.close()


# ----------------------------------------------------------------------- wrappers
def test_history_and_subset_wrappers():
    env = HistoryWrapper(StubEnv(), max_history=4)
    state = env.reset()
    assert env.observation_space.contains({k: np.asarray(v, np.float32) for k, v in state.items()})
    assert state['test1d'].shape == (4, 5)
    first = state['test1d'][0].copy()
    state, _, _, _ = env.step(env.action_space.sample())
    assert np.array_equal(state['test1d'][1], first)      # newest first
    sub = SubSetWrapper(StubEnv(), ['test2d'])
    assert list(sub.reset()) == ['test2d'] and list(sub.step(None)[0]) == ['test2d']
    assert list(sub.observation_space.spaces) == ['test2d']


@pytest.mark.parametrize('monitor_cls', [StrictMonitor, Monitor])
def test_monitor_csv(monitor_cls):
    with tempfile.TemporaryDirectory() as tmp:
        env = monitor_cls(StubEnv(), os.path.join(tmp, 'log'), chunk_size=1)
        for _ in range(3):
            env.reset()
            done = False
            while not done:
                _, _, done, info = env.step(None)
            assert info['episode']['l'] == 10
        env.close()
        frame = pd.read_csv(os.path.join(tmp, 'log.mon.csv'))
        assert len(frame) == 3 and {'r', 'l', 't'} <= set(frame.columns)
        assert list(frame.columns) == sorted(frame.columns)
        assert env.get_episode_lengths() == [10, 10, 10]


def test_strict_monitor_reset_rules():
    env = StrictMonitor(StubEnv(), None)
    with pytest.raises(RuntimeError):
        env.step(None)
    env.reset()
    with pytest.raises(RuntimeError):
        env.reset()


def test_gym_ids_registered():
    from custom_envs_b200.compat import gym_standin
    if not hasattr(gym, '__standin__'):
        pytest.skip('real gym present')
    assert 'MultiOptLRs-v0' in gym_standin._REGISTRY


def test_stage_to_device_chunks_cover_the_vector():
    """The staging helper of the host-facing VecEnv (numpy -> pinned -> device in chunks handled by
    the library's own threads) on CPU tensors: ragged last chunk, single chunk, one element."""
    import torch
    from custom_envs_b200.vectorize.optvecenv import stage_to_device
    for count, chunk in [(100003, 4096), (4096, 4096), (1, 8), (12345, 1 << 23)]:
        src = np.random.RandomState(count).rand(count).astype(np.float32)
        host, dev = torch.zeros(count), torch.zeros(count)
        stage_to_device(src, host, dev, chunk=chunk)
        assert np.array_equal(dev.numpy(), src) and np.array_equal(host.numpy(), src)


def test_headers_are_plain_c_and_a_c_client_resolves_every_entry_point(tmp_path):
    """include/*.h compile as C99 (-pedantic -Werror) and a dlopen client written in C finds every
    declared symbol; the calls that need no GPU return what the headers document."""
    import shutil
    import subprocess
    if shutil.which('gcc') is None:
        pytest.skip('no C compiler')
    here = os.path.dirname(os.path.abspath(__file__))
    exe = str(tmp_path / 'c_abi_smoke')
    subprocess.run(['gcc', '-std=c99', '-pedantic', '-Wall', '-Werror', '-I', os.path.join(here, '..', 'include'),
                    os.path.join(here, 'c_abi_smoke.c'), '-o', exe, '-ldl'], check=True)
    out = subprocess.run([exe, _lib.LIB_PATH], check=True, capture_output=True, text=True)
    assert 'c abi ok' in out.stdout


def test_device_dataset_container_follows_the_batching_rules():
    """DeviceDataSet (what load_data(..., device=) returns) on CPU tensors: ceil(N/B) batches, ragged
    last batch, class count inferred from the label ranks (dataset/inmemorydataset.py:11-28)."""
    import torch
    from custom_envs_b200.dataset import DeviceDataSet
    feats, ranks = torch.arange(150 * 4, dtype=torch.float32).reshape(150, 4), torch.arange(150, dtype=torch.int32) % 3
    data = DeviceDataSet(feats, ranks, batch_size=32)
    assert len(data) == 5 and data.num_classes == 3 and data.on_device
    assert data.feature_shape == (4,) and data.target_shape == (3,)
    assert [len(batch.features) for batch in data] == [32, 32, 32, 32, 22]
    assert torch.equal(data[4].labels, ranks[128:])
    assert len(DeviceDataSet(feats, ranks)) == 1                      # batch_size=None: the whole set
    with pytest.raises(AssertionError):
        DeviceDataSet(feats, ranks[:10])


def test_policy_actions_chunks_and_clips():
    """The chunked policy evaluation of device_rollout on CPU tensors: every row is visited once
    whatever the chunk size, actions are clipped to MultiOptLRs' Box [-4, 6]."""
    import torch
    from custom_envs_b200.vectorize.device_rollout import SharedMlpPolicy, policy_actions
    obs = torch.linspace(-30, 30, 101 * 15).reshape(101, 15)
    calls = []

    def act(chunk):
        calls.append(len(chunk))
        return chunk.sum(dim=1)

    want = obs.sum(dim=1).clamp(-4.0, 6.0)
    for chunk in (1, 7, 101, 1000):
        calls.clear()
        out = policy_actions(act, obs, torch.empty(101), row_chunk=chunk)
        assert torch.equal(out, want) and sum(calls) == 101 and max(calls) <= chunk
    torch.manual_seed(0)
    policy = SharedMlpPolicy(15)
    mean, value = policy(obs)
    assert mean.shape == value.shape == (101,)
    gen = torch.Generator().manual_seed(1)
    assert policy.act(obs, generator=gen).shape == (101,)


def test_fuse_key_is_the_data_content_not_its_shape_and_sums():
    """Envs are fused onto one HBM replica of a data set only if their arrays are EQUAL: the same
    rows in another order (equal shape, equal sums) must not share the first env's arrays."""
    from custom_envs_b200.compat import make
    import custom_envs  # noqa: F401  (registers the env ids)
    from custom_envs_b200.dataset.inmemorydataset import InMemoryDataSet
    rng = np.random.RandomState(0)
    feats = rng.uniform(size=(40, 4))
    targs = np.eye(3)[np.arange(40) % 3]
    data_sets = [InMemoryDataSet(feats, targs, 8), InMemoryDataSet(feats[::-1].copy(), targs[::-1].copy(), 8),
                 InMemoryDataSet(feats.copy(), targs.copy(), 8)]
    envs = [make('MultiOptLRs-v0', problem='nn', problem_kwargs=dict(layers=(), data_set=data))
            for data in data_sets + data_sets[:1]]
    keys = [getattr(env, 'unwrapped', env).fuse_key() for env in envs]
    assert keys[0] != keys[1]                      # permuted rows: not the same data
    assert keys[0] == keys[2] == keys[3]           # equal content / the same object: fusable
    assert hasattr(data_sets[0], '_b2e_digest')    # hashed once per data-set object

"""Data front-end: the oracle against vectors recorded from the reference's own
utils_image / utils_math / utils_common functions (tests/golden/gen_data_golden.py), and the
CUDA kernels (through the C ABI in include/b200data.h) against both.  Bit-exact throughout:
byte gathers, float64 arithmetic with IEEE division, integer ranks."""
import os

import numpy as np
import pytest

from oracle import data_oracle as orc

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'data_frontend.npz'))
RESIZE_CASES = 6
LABEL_CASES = {0: None, 1: 12, 2: None, 3: 3}


# ------------------------------------------------------------------ oracle vs reference vectors
@pytest.mark.parametrize('case', range(RESIZE_CASES))
def test_oracle_resize_matches_pillow(case):
    got = orc.resize_nearest(GOLDEN['resize%d_in' % case], tuple(GOLDEN['resize%d_shape' % case]))
    assert np.array_equal(got, GOLDEN['resize%d_out' % case])


def test_oracle_nearest_table_against_pillow_for_many_sizes():
    from PIL import Image
    for src in (5, 7, 28, 31, 32, 100):
        ramp = np.arange(src, dtype=np.uint8)[None, :].repeat(2, axis=0)
        for dst in (1, 2, 3, 7, 10, 13, src):
            want = np.asarray(Image.fromarray(ramp).resize((dst, 2), 0))[0]
            assert np.array_equal(orc.pillow_nearest_table(src, dst), want), (src, dst)


@pytest.mark.parametrize('name', ['float', 'float32', 'int'])
def test_oracle_normalize_matches_reference(name):
    got = orc.normalize(GOLDEN[name + '_in'])
    assert got.dtype == GOLDEN[name + '_out'].dtype and np.array_equal(got, GOLDEN[name + '_out'])


def test_oracle_mnist_branch_matches_reference():
    got = orc.image_features(GOLDEN['mnist_in'].reshape(-1, 28, 28))
    assert np.array_equal(got, GOLDEN['mnist_features'])
    assert np.all(got[:, 0] == 0.0)           # source pixel 58 = (2, 2) -> feature 0 is constant: 0 / 1e-8


@pytest.mark.parametrize('case', sorted(LABEL_CASES))
def test_oracle_onehot_matches_reference(case):
    onehot, num = orc.to_onehot(GOLDEN['labels%d_in' % case], LABEL_CASES[case])
    assert num == int(GOLDEN['labels%d_num' % case])
    assert np.array_equal(onehot, GOLDEN['labels%d_onehot' % case])


def test_host_mirror_table_equals_oracle_table():
    from custom_envs_b200.data.device_frontend import pillow_nearest_table
    for src, dst in [(28, 7), (28, 10), (32, 7), (9, 4), (13, 5), (100, 33)]:
        assert np.array_equal(pillow_nearest_table(src, dst), orc.pillow_nearest_table(src, dst))


def test_frontend_refuses_cpu_tensors():
    import torch
    from custom_envs_b200 import _lib
    from custom_envs_b200.data import device_frontend as dev
    if torch.cuda.is_available():
        pytest.skip('GPU present: the CUDA path is taken')
    with pytest.raises((_lib.B200EnvError, RuntimeError, AssertionError)):
        dev.normalize(np.zeros((4, 3)), device='cuda:0')


# ------------------------------------------------------------------ CUDA vs oracle / golden
@pytest.mark.gpu
@pytest.mark.parametrize('case', range(RESIZE_CASES))
def test_device_resize_matches_pillow(case):
    from custom_envs_b200.data import device_frontend as dev
    got = dev.resize_nearest(GOLDEN['resize%d_in' % case], tuple(GOLDEN['resize%d_shape' % case]))
    assert np.array_equal(got.cpu().numpy(), GOLDEN['resize%d_out' % case])


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['float', 'float32', 'int'])
def test_device_normalize_matches_reference(name):
    import torch
    from custom_envs_b200.data import device_frontend as dev
    got = dev.normalize(GOLDEN[name + '_in'], out_dtype=torch.float64)
    assert np.array_equal(got.cpu().numpy(), GOLDEN[name + '_out'])
    got32 = dev.normalize(GOLDEN[name + '_in'], out_dtype=torch.float32)
    assert np.array_equal(got32.cpu().numpy(), GOLDEN[name + '_out'].astype(np.float32))


@pytest.mark.gpu
def test_device_mnist_branch_matches_reference():
    from custom_envs_b200.data import device_frontend as dev
    labels = np.arange(64) % 10
    features, ranks, num = dev.image_dataset(GOLDEN['mnist_in'], labels)
    assert num == 10 and np.array_equal(ranks.cpu().numpy(), labels)
    assert np.array_equal(features.cpu().numpy(), GOLDEN['mnist_features'].astype(np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize('case', sorted(LABEL_CASES))
def test_device_onehot_matches_reference(case):
    from custom_envs_b200.data import device_frontend as dev
    onehot, num = dev.to_onehot(GOLDEN['labels%d_in' % case], LABEL_CASES[case])
    assert num == int(GOLDEN['labels%d_num' % case])
    assert np.array_equal(onehot.cpu().numpy(), GOLDEN['labels%d_onehot' % case])


@pytest.mark.gpu
def test_device_onehot_raises_where_numpy_raises():
    from custom_envs_b200 import _lib
    from custom_envs_b200.data import device_frontend as dev
    with pytest.raises(IndexError):
        orc.to_onehot(np.array([0, 1, 2, 3]), 3)
    with pytest.raises(_lib.B200EnvError):
        dev.to_onehot(np.array([0, 1, 2, 3]), 3)
    with pytest.raises(_lib.B200EnvError):
        dev.label_ranks(np.array([0, 70000]))


@pytest.mark.gpu
@pytest.mark.parametrize('rows,cols,dtype', [(1, 1, np.float64), (3, 784, np.uint8), (60000, 49, np.uint8),
                                             (20011, 784, np.float32), (245057, 3, np.float64),
                                             (1000, 256, np.int32), (777, 3000, np.float64)])
def test_device_normalize_shapes_against_oracle(rows, cols, dtype):
    """Ragged and full-size tables (60000 x 49 is the reference's MNIST shape, 245057 x 3 its
    skin table), every kernel path: narrow rows, wide rows, one row, padded output rows."""
    import torch
    from custom_envs_b200.data import device_frontend as dev
    rng = np.random.RandomState(rows + cols)
    if np.dtype(dtype).kind == 'f':
        table = (rng.normal(size=(rows, cols)) * 50).astype(dtype)
    else:
        table = rng.randint(0, 256, size=(rows, cols)).astype(dtype)
    want = orc.normalize(table)
    mins, maxes = dev.column_minmax(table)
    assert np.array_equal(mins.cpu().numpy(), table.min(0).astype(np.float64))
    assert np.array_equal(maxes.cpu().numpy(), table.max(0).astype(np.float64))
    padded = -(-cols // 4) * 4
    out = torch.full((rows, padded), -7.0, dtype=torch.float32, device='cuda:0')
    dev.normalize(table, out=out)
    got = out.cpu().numpy()
    assert np.array_equal(got[:, :cols], want.astype(np.float32))
    assert np.all(got[:, cols:] == -7.0)
    # size-independent properties: range and idempotence up to the 1e-8 guard
    assert got[:, :cols].min() >= 0.0 and got[:, :cols].max() <= 1.0


@pytest.mark.gpu
def test_device_dataset_feeds_the_env_like_the_host_dataset():
    """Raw bytes -> device front-end -> BatchedOptEnv gives the same first observation and step
    as the host-prepared (oracle) features uploaded as float32."""
    import torch
    from custom_envs_b200.batched_env import BatchedOptEnv, ProblemSpec
    from custom_envs_b200.data import device_frontend as dev
    rng = np.random.RandomState(5)
    images = rng.randint(0, 256, size=(300, 784)).astype(np.uint8)
    labels = rng.randint(0, 10, 300)
    features, ranks, num = dev.image_dataset(images, labels)
    host_features = orc.image_features(images.reshape(-1, 28, 28)).astype(np.float32)
    outs = []
    for feats, targs in ((features, ranks), (host_features, labels.astype(np.int32))):
        env = BatchedOptEnv(ProblemSpec('softmax', 49, (), num), feats, targs, 4, batch_size=32,
                            max_batches=20, seeds=[0, 1, 2, 3], init_seed=3)
        obs0 = env.reset().clone()
        actions = torch.full((env.num_rows,), 1.5, device=env.device)
        obs1, reward, _, _ = env.step(actions)
        outs.append((obs0.cpu().numpy(), obs1.cpu().numpy(), reward.cpu().numpy()))
        env.close()
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


@pytest.mark.gpu
def test_load_data_on_device_equals_host_preparation():
    """load_data(name, device=...) (raw table -> CUDA front-end) against the host path of the
    same loader, and an OptimizeNN env built on the device data set."""
    from custom_envs_b200.data import load_data
    host = load_data('iris', batch_size=32)
    dev_set = load_data('iris', batch_size=32, device='cuda:0')
    assert len(dev_set) == len(host) == 5 and dev_set.num_classes == 3
    assert np.array_equal(dev_set.features.cpu().numpy(), host.features.astype(np.float32))
    assert np.array_equal(dev_set.targets.cpu().numpy(), host.targets.argmax(axis=1))
    from custom_envs_b200.envs.multioptlrs import MultiOptLRs
    outs = []
    action = {'parameter-%d' % i: np.array([1.0 + 0.1 * i], np.float32) for i in range(15)}
    for data_set in (host, dev_set):
        env = MultiOptLRs(problem='nn', max_batches=10, problem_kwargs=dict(layers=(), data_set=data_set))
        env.seed(3)
        env.reset()
        env.model.set_parameters(np.linspace(-0.5, 0.5, 15))
        for _ in range(2):
            state, reward, _, info = env.step(action)
        outs.append((np.stack([state[key] for key in sorted(state)]), reward, info['batch_loss']))
        env.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and outs[0][1:] == outs[1][1:]


@pytest.mark.gpu
def test_device_shuffle_is_numpys_randomstate_shuffle_bit_for_bit():
    """the per-env epoch permutation, `np.random.RandomState(seed).shuffle(np.arange(N))`
    (utils/utils_common.py:12-23 under utils/utils_math.py:10-22), generated on the device: MT19937 seeding, tempering,
    the masked rejection sampling of `random_interval` and the Fisher-Yates order restated.  Integer seeds, generators
    in an arbitrary state (classic gym's hash-seeded RandomState, a generator that has already been drawn from), the
    sizes of the BASELINE data sets, tiny and degenerate lengths, more generators than one warp."""
    import torch
    from custom_envs_b200.batched_env import env_permutations, env_permutations_device
    from custom_envs_b200.compat import gym_standin
    for num_rows, seeds in [(150, list(range(70))), (60000, [0, 1, 2, 12345, 2 ** 32 - 1]), (1, [3]), (2, [4, 5]),
                            (623, [7]), (625, [8]), (4097, list(range(1000, 1040)))]:
        want = env_permutations(num_rows, seeds)
        got = env_permutations_device(num_rows, seeds, 'cuda:0').cpu().numpy()
        assert np.array_equal(got, want), (num_rows, seeds[:3])
    # generators in arbitrary states: gym's np_random(seed) (init_by_array of a hashed seed), and one that was used before
    gens = [gym_standin.np_random(s)[0] for s in range(5)]
    used = np.random.RandomState(99)
    used.uniform(size=1000)                                   # position inside the 624-word block
    gens.append(used)
    want = env_permutations(3000, gens)                       # copies the state, like use_random_state
    got = env_permutations_device(3000, gens, 'cuda:0').cpu().numpy()
    assert np.array_equal(got, want)
    assert np.array_equal(env_permutations(3000, gens), want)  # ... and leaves the generators where they were

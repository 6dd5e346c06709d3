"""CPU tests that pin the oracle.

1. Replay of the golden fixtures in tests/golden/*.npz.  Those were produced by
   tests/golden/gen_golden.py, which runs the reference's own MultiOptLRs /
   MultiOptimize / History / utils_env / InMemoryDataSet / OptVecEnv code (imported by
   path from /root/reference in the build container).
2. The forward/backward restatement against torch.autograd (float64).
3. Re-expression of the reference's layout tests (tests/utils/test_utils_common.py:68-76,
   104-114; tests/utils/test_utils_math.py:15-24; tests/dataset/test_inmemorydataset.py:27-40).
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import optenv_oracle as orc

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden', 'opt*.npz')))


def load_fixture(path):
    fix = dict(np.load(path, allow_pickle=False))
    kwargs = dict(zip([str(k) for k in fix['env_kwargs_keys']],
                      [int(v) for v in fix['env_kwargs_vals']]))
    spec = orc.ProblemSpec(str(fix['problem_kind']), int(fix['num_features']),
                           tuple(int(h) for h in fix['hidden']), int(fix['num_outputs']))
    if str(fix['env_kind']) == 'optlrs':
        cfg = orc.EnvConfig.multioptlrs(**kwargs)
    else:
        cfg = orc.EnvConfig.multioptimize(**kwargs)
    return fix, spec, cfg


def build_oracle(fix, spec, cfg, **extra):
    num_envs = int(fix['num_envs'])
    if spec.kind == 'func':
        return orc.BatchedOptEnvOracle(spec, None, None, num_envs, config=cfg, **extra)
    batch = int(fix['batch_size'])
    return orc.BatchedOptEnvOracle(
        spec, fix['feats'], fix['labels'], num_envs, batch_size=None if batch < 0 else batch,
        config=cfg, perms=fix['perms'], init_orders=fix['init_orders'], **extra)


def test_golden_fixtures_present():
    assert len(GOLDEN) >= 25


@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_reference_run(path):
    fix, spec, cfg = load_fixture(path)
    env = build_oracle(fix, spec, cfg)
    vec = orc.OptVecEnvOracle(env)
    num_envs, num_params = env.num_envs, env.num_params
    reset_no = np.zeros(num_envs, int)
    batch_no = np.zeros(num_envs, int)

    def next_params():
        last = fix['reset_params'].shape[1] - 1
        return np.stack([fix['reset_params'][e, min(reset_no[e], last)]
                         for e in range(num_envs)])

    def check_batches():
        if spec.kind == 'func':
            return
        idx, cnt = env.stream.current()
        for e in range(num_envs):
            want = fix['batches'][e, batch_no[e]]
            want = want[want >= 0]
            assert cnt[e] == len(want)
            assert np.array_equal(idx[e, :cnt[e]], want)     # bit-exact index stream

    states = vec.reset(init_params=next_params())
    reset_no += 1
    assert np.array_equal(states, fix['reset_states'])
    check_batches()
    for t in range(fix['actions'].shape[0]):
        states, rewards, dones, info = vec.step(fix['actions'][t], reset_params=next_params())
        env_done = dones[::num_params]
        reset_no += env_done
        if cfg.env == 'optlrs':
            batch_no += 1
        batch_no += env_done
        check_batches()
        assert np.array_equal(dones, fix['dones'][t])
        np.testing.assert_allclose(states, fix['states'][t], rtol=1e-12, atol=0)
        np.testing.assert_allclose(rewards, fix['rewards'][t], rtol=1e-12, atol=0)
        for key in orc.INFO_KEYS:
            np.testing.assert_allclose(info[key], fix['info_' + key][t], rtol=1e-10,
                                       atol=0, equal_nan=True, err_msg=key)
        np.testing.assert_allclose(info['episode_r'], fix['info_episode_r'][t], rtol=1e-12)
        assert np.array_equal(info['episode_l'], fix['info_episode_l'][t])


def _torch_loss(spec, theta, feats, targs):
    params, start = [], 0
    for shape in spec.shapes:
        n = int(np.prod(shape))
        params.append(theta[start:start + n].reshape(shape))
        start += n
    cur = feats
    nlayers = len(params) // 2
    for li in range(nlayers):
        cur = cur @ params[2 * li] + params[2 * li + 1]
        if li < nlayers - 1:
            cur = torch.relu(cur)
    if spec.kind == 'softmax':
        return torch.nn.functional.cross_entropy(cur, targs, reduction='none')
    return 0.5 * ((cur - targs) ** 2).sum(dim=1)


@pytest.mark.parametrize('spec', [
    orc.ProblemSpec('softmax', 4, (), 3), orc.ProblemSpec('softmax', 20, (7,), 5),
    orc.ProblemSpec('softmax', 9, (6, 4), 3), orc.ProblemSpec('linreg', 4, (), 1),
    orc.ProblemSpec('linreg', 5, (), 3)])
def test_loss_and_grad_vs_autograd(spec):
    rng = np.random.RandomState(0)
    num_envs, batch = 3, 11
    theta = rng.normal(size=(num_envs, spec.size)).astype(np.float32)
    feats = rng.uniform(size=(num_envs, batch, spec.num_features)).astype(np.float32)
    if spec.kind == 'softmax':
        targs = rng.randint(0, spec.num_outputs, size=(num_envs, batch))
    else:
        targs = rng.normal(size=(num_envs, batch, spec.num_outputs)).astype(np.float32)
    cnt = np.array([batch, batch - 4, 1])
    mask = np.arange(batch)[None, :] < cnt[:, None]
    grad, loss = orc.loss_and_grad(spec, theta, feats, targs, mask)
    for e in range(num_envs):
        th = torch.tensor(theta[e], dtype=torch.float64, requires_grad=True)
        x = torch.tensor(feats[e, :cnt[e]], dtype=torch.float64)
        y = torch.tensor(targs[e, :cnt[e]])
        y = y.long() if spec.kind == 'softmax' else y.double()
        per_sample = _torch_loss(spec, th, x, y)
        per_sample.sum().backward()                 # tf.gradients of a vector = SUM
        np.testing.assert_allclose(loss[e], per_sample.mean().item(), rtol=1e-12)
        np.testing.assert_allclose(grad[e], th.grad.numpy(), rtol=1e-10, atol=1e-13)


def test_rosenbrock():
    spec = orc.ProblemSpec('func', 0, (), 0)
    theta = np.array([[-1.9, 2.0], [1.0, 1.0]])
    grad, loss = orc.loss_and_grad(spec, theta, None, None, None)
    th = torch.tensor(theta, dtype=torch.float64, requires_grad=True)
    val = 100 * (th[:, 1] - th[:, 0] ** 2) ** 2 + (1 - th[:, 0]) ** 2
    val.sum().backward()
    np.testing.assert_allclose(loss, val.detach().numpy(), rtol=1e-14)
    np.testing.assert_allclose(grad, th.grad.numpy(), rtol=1e-12, atol=1e-12)


def test_lexicographic_rows():
    perm = orc.lexicographic_rows(112)
    names = ['parameter-%d' % i for i in range(112)]
    assert [names[i] for i in perm] == sorted(names)
    assert list(perm[:5]) == [0, 1, 10, 100, 101]


def test_index_stream_layout():
    # tests/dataset/test_inmemorydataset.py:27-40: len = ceil(N/B), ragged last batch
    perm = orc.env_permutation(10, 3)
    stream = orc.IndexStream(10, 4, perm[None])
    everyone = np.ones(1, bool)
    stream.reset(everyone)
    seen = []
    for want in (4, 4, 2):
        idx, cnt = stream.current()
        assert cnt[0] == want
        seen += list(idx[0, :cnt[0]])
        stream.advance(everyone)
    assert sorted(seen) == list(range(10))
    assert seen == list(perm)                       # identity order permuted once
    idx, cnt = stream.current()                     # epoch 2: the SAME permutation again
    assert list(idx[0, :4]) == list(perm[perm][:4])


def test_env_permutation_is_fresh_randomstate_shuffle():
    # tests/utils/test_utils_math.py:15-24: draws in the context == fresh RandomState(seed)
    want = np.arange(25)
    np.random.RandomState(5).shuffle(want)
    assert np.array_equal(orc.env_permutation(25, 5), want)


def test_observation_layout_newest_first():
    # tests/utils/test_utils_common.py:68-76,104-114 re-expressed on the oracle's rings
    spec = orc.ProblemSpec('func', 0, (), 0)
    env = orc.BatchedOptEnvOracle(spec, None, None, 1,
                                  config=orc.EnvConfig.multioptlrs(max_history=3))
    obs = env.reset()
    assert obs.shape == (1, 2, 9) and np.all(obs == -1.0)
    marks = []
    for t in range(4):
        obs, _, _, info = env.step(np.full((1, 2), -3.0, np.float32))
        marks.append(info['adjusted_loss'][0])
        want = np.clip(marks[::-1][:3], -100, 100) - 1
        np.testing.assert_allclose(obs[0, 0, 3:3 + len(want)], want)
        np.testing.assert_allclose(obs[0, 1, 3:3 + len(want)], want)

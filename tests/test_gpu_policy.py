"""GPU tests of the shared per-agent policy kernel (include/b200policy.h) and of the ring-only env
step that feeds it (b2e_step with obs_out = NULL), SURVEY 8f.2.

The policy is the caller's model (stable-baselines MlpPolicy evaluated on every agent row,
reference run_multiagent_exp_single.py:37-49), so it is held to a torch evaluation of the same
network that rounds to bf16 where the kernel does -- not to the env step's 1e-5 bar.  The ring-only
step IS the env step: its state, rewards, done flags and info rows must equal the dense step's."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from custom_envs_b200.batched_env import BatchedOptEnv, ProblemSpec, env_permutations
    from custom_envs_b200.vectorize.device_policy import (DevicePolicy, device_policy_rollout,
                                                          reference_actions)
    from custom_envs_b200.vectorize.device_rollout import SharedMlpPolicy


def make_policy(obs_dim, seed, scale=1.0):
    torch.manual_seed(seed)
    policy = SharedMlpPolicy(obs_dim).cuda()
    with torch.no_grad():
        for p in policy.pi.parameters():
            p.mul_(scale)
        for m in policy.pi.modules():
            if isinstance(m, torch.nn.Linear):
                m.bias.uniform_(-0.3, 0.3)
    return policy


@pytest.mark.parametrize('tanh_mode', [0, 1, 2])
@pytest.mark.parametrize('obs_dim,rows', [(15, 128 * 9), (15, 100003), (9, 4321), (15, 1), (3, 257)])
def test_dense_policy_matches_bf16_torch(obs_dim, rows, tanh_mode):
    policy = make_policy(obs_dim, seed=obs_dim + rows, scale=2.0)
    gen = torch.Generator(device='cuda').manual_seed(rows)
    obs = torch.randn(rows, obs_dim, device='cuda', generator=gen)
    obs[::7] *= 10.0
    obs[::13, 0] = 99.0          # the clip bounds of the observation (multioptlrs.py:99)
    obs[::17, -1] = -101.0
    dev = DevicePolicy.from_torch(policy.pi, log_std=None, tanh_mode=tanh_mode, low=-1e9, high=1e9)
    got = dev.act(obs)
    want = reference_actions(policy.pi, obs, low=-1e9, high=1e9)
    exact = reference_actions(policy.pi, obs, low=-1e9, high=1e9, bf16=False)
    torch.cuda.synchronize()
    err = (got - want).abs().max().item()
    # tanh.approx.f32 is good to ~2^-11 relative, the packed bf16 variant to bf16 resolution; the head sums 64 terms
    scale = policy.pi[-1].weight.abs().sum().item()
    assert err <= (2e-3 if tanh_mode == 0 else 1.5e-2) * max(1.0, scale), (err, scale)
    assert (got - exact).abs().max().item() <= 0.1 * max(1.0, scale)          # bf16 operands against fp32
    # clipping to the action Box
    dev2 = DevicePolicy.from_torch(policy.pi, tanh_mode=tanh_mode, low=-0.05, high=0.07)
    assert torch.equal(dev2.act(obs), got.clamp(-0.05, 0.07))
    dev.close()
    dev2.close()


def test_unaligned_observation_matrix_takes_the_plain_load_path():
    policy = make_policy(15, seed=5)
    base = torch.randn(1000 * 15 + 1, device='cuda')
    obs = base[1:].view(1000, 15)                    # 4-byte aligned only: no bulk copies
    dev = DevicePolicy.from_torch(policy.pi, low=-1e9, high=1e9)
    got = dev.act(obs)
    want = dev.act(obs.clone())
    assert torch.equal(got, want)


def test_gaussian_noise_is_per_row_and_reproducible():
    policy = make_policy(15, seed=3)
    obs = torch.zeros(1 << 18, 15, device='cuda')
    quiet = DevicePolicy.from_torch(policy.pi, low=-1e9, high=1e9)
    noisy = DevicePolicy.from_torch(policy.pi, log_std=torch.tensor(-1.0), low=-1e9, high=1e9)
    base = quiet.act(obs)
    a, b, c = noisy.act(obs, seed=11), noisy.act(obs, seed=11), noisy.act(obs, seed=12)
    assert torch.equal(a, b) and not torch.equal(a, c)
    z = (a - base) / np.exp(-1.0)
    assert abs(z.mean().item()) < 0.01 and abs(z.std().item() - 1.0) < 0.01
    assert abs((z ** 4).mean().item() - 3.0) < 0.15               # Gaussian kurtosis
    assert abs(torch.corrcoef(torch.stack([z[:-1], z[1:]]))[0, 1].item()) < 0.01


def small_env(num_envs=3, max_batches=6, row_order='lexicographic', materialize_obs=True, seed=0, max_history=5,
              hidden=(64,)):
    """The BASELINE config-4 problem shape (MLP 784-64-10: the tcgen05 pipeline) on a small data set; other
    ``hidden`` tuples select the other large-problem pipelines (() = config 3's resident-minibatch eval kernel,
    two layers = the generic dense stack)."""
    rng = np.random.RandomState(seed)
    rows = 320
    feats = rng.uniform(size=(rows, 784)).astype(np.float32)
    labels = rng.randint(0, 10, rows).astype(np.int32)
    env = BatchedOptEnv(ProblemSpec('softmax', 784, tuple(hidden), 10), feats, labels, num_envs, batch_size=32,
                        max_batches=max_batches, max_history=max_history, row_order=row_order,
                        perms=env_permutations(rows, list(range(num_envs))), init_seed=9,
                        materialize_obs=materialize_obs)
    env.reset()
    return env


SUMMED_TWICE = [6, 7, 12, 13]           # states_mean, states_sum, adjusted_grad, grad_diff (multioptlrs.py:119-126)


def check_info(got, want, msg):
    got, want = got.cpu().numpy(), want.cpu().numpy()
    rest = [c for c in range(got.shape[1]) if c not in SUMMED_TWICE]
    np.testing.assert_array_equal(got[:, rest], want[:, rest], err_msg=msg)
    a, b = got[:, SUMMED_TWICE], want[:, SUMMED_TWICE]
    # sums holding a nan_to_num(x/0) = FLT_MAX term (zero-initialised biases) overflow in one grouping of the fp32
    # partial sums and not in the other: outside the parity domain; both sides must agree that they are huge
    huge = ~(np.isfinite(a) & (np.abs(a) < 1e30)) | ~(np.isfinite(b) & (np.abs(b) < 1e30))
    assert np.all((np.abs(a[huge]) > 1e30) & (np.abs(b[huge]) > 1e30)), msg
    np.testing.assert_allclose(a[~huge], b[~huge], rtol=2e-6, atol=0, err_msg=msg)


@pytest.mark.parametrize('row_order,depth', [('lexicographic', 5), ('natural', 5), ('lexicographic', 3), ('natural', 1)])
def test_ring_front_end_equals_dense_front_end(row_order, depth):
    """act_env reads the rings, act reads the observation rows the same step wrote: identical actions,
    from the reset observation (all -1), through a filling history, to a full one."""
    env = small_env(row_order=row_order, max_history=depth)          # depth != 5: the run-time-depth ring front end
    assert env.obs_dim == 3 * depth
    policy = make_policy(env.obs_dim, seed=1, scale=1.5)
    dev = DevicePolicy.from_torch(policy.pi)
    gen = torch.Generator(device='cuda').manual_seed(0)
    for t in range(8):                                       # max_batches = 6: an auto-reset happens inside
        dense = dev.act(env.obs)
        ring = dev.act_env(env)
        assert torch.equal(dense, ring), t
        want = reference_actions(policy.pi, env.obs)
        assert (dense - want).abs().max().item() < 5e-3
        env.step(torch.rand(env.num_rows, device='cuda', generator=gen) * 2.0)
    dev.close()
    env.close()


@pytest.mark.parametrize('hidden', [(64,), (), (24, 8)], ids=['tcgen05', 'resident-minibatch', 'generic-stack'])
def test_ring_only_step_is_the_same_env_step(hidden):
    """Two identical env batches, one stepping with observation rows, one ring-only (one of them without
    an observation matrix at all): parameters, rings, rewards, done flags and all 16 info columns agree
    step by step, across an episode end (the four statistics that the two paths sum in different fp32
    groupings -- states_mean / states_sum rebuilt from per-slot sums, adjusted_grad, grad_diff -- to 2e-6
    relative); a dense step afterwards returns identical observation rows."""
    dense, ring = small_env(hidden=hidden), small_env(materialize_obs=False, hidden=hidden)
    assert ring.obs is None
    gen = torch.Generator(device='cuda').manual_seed(4)
    for t in range(9):
        actions = torch.rand(dense.num_rows, device='cuda', generator=gen) * 2.5
        _, rew_d, done_d, info_d = dense.step(actions)
        obs_r, rew_r, done_r, info_r = ring.step(actions)
        assert obs_r is None
        assert torch.equal(rew_d, rew_r) and torch.equal(done_d, done_r), t
        check_info(info_r, info_d, 'step %d' % t)
        for name in ('params', 'grad_prev', 'adj_weights', 'adj_grads', 'adj_losses', 'step', 'cursor'):
            assert torch.equal(dense.get_state(name), ring.get_state(name)), (t, name)
    # the env that never wrote a row can still produce them
    ring.obs = torch.empty_like(dense.obs)
    actions = torch.rand(dense.num_rows, device='cuda', generator=gen)
    obs_d = dense.step(actions)[0]
    obs_r = ring.step(actions)[0]
    assert torch.equal(obs_d, obs_r)
    check_info(ring.info, dense.info, 'dense step after ring-only steps')
    # set_state of the rings refreshes the per-slot sums the ring-only statistics use
    ring.set_state('adj_weights', dense.get_state('adj_weights') * 2.0)
    dense.set_state('adj_weights', dense.get_state('adj_weights') * 2.0)
    actions = torch.rand(dense.num_rows, device='cuda', generator=gen)
    dense.step(actions)
    ring.step(actions, ring_only=True)
    check_info(ring.info, dense.info, 'after set_state')
    dense.close()
    ring.close()


def test_ring_only_needs_the_large_problem_pipeline():
    from custom_envs_b200._lib import B200EnvError
    rng = np.random.RandomState(0)
    feats = rng.uniform(size=(150, 4)).astype(np.float32)
    labels = (np.arange(150) % 3).astype(np.int32)
    env = BatchedOptEnv(ProblemSpec('softmax', 4, (), 3), feats, labels, 4, perms=env_permutations(150, [0, 1, 2, 3]))
    env.reset()
    with pytest.raises(B200EnvError, match='obs_out = NULL'):
        env.step(torch.zeros(env.num_rows, device='cuda'), ring_only=True)
    dev = DevicePolicy(15)
    dev.set_weights(*[torch.zeros(s, device='cuda') for s in ((64, 15), (64,), (64, 64), (64,), (1, 64), (1,))])
    with pytest.raises(B200EnvError, match='rings'):
        dev.act_env(env)
    env.close()


def test_policy_rollout_on_the_rings_equals_rollout_on_observation_rows():
    """device_policy_rollout: the closed loop policy -> step -> policy ... gives the same returns whether the
    policy reads observation rows or the rings (deterministic policy: the two loops are the same computation)."""
    policy = make_policy(15, seed=2, scale=1.5)
    results = []
    for ring_only in (False, True):
        env = small_env(num_envs=4, max_batches=5, materialize_obs=not ring_only)
        dev = DevicePolicy.from_torch(policy.pi)
        returns, finished = device_policy_rollout(env, dev, 7, ring_only=ring_only)
        results.append((returns.cpu().numpy(), finished, env.get_state('params').cpu().numpy()))
        dev.close()
        env.close()
    assert results[0][1] == results[1][1] == 4
    assert np.array_equal(results[0][0], results[1][0]) and np.array_equal(results[0][2], results[1][2])


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_env_and_policy_on_a_device_that_is_not_current():
    """Every entry point runs on the handle's own device whatever the caller's current device is: an env
    batch and a policy built with device='cuda:1' while cuda:0 is current give the trajectories of the same objects
    on cuda:0, bit for bit."""
    results = []
    for dev in ('cuda:0', 'cuda:1'):
        torch.cuda.set_device(0)
        rng = np.random.RandomState(0)
        feats = rng.uniform(size=(320, 784)).astype(np.float32)
        labels = rng.randint(0, 10, 320).astype(np.int32)
        env = BatchedOptEnv(ProblemSpec('softmax', 784, (64,), 10), feats, labels, 3, batch_size=32, max_batches=5,
                            perms=env_permutations(320, [0, 1, 2]), init_seed=9, device=dev)
        assert torch.cuda.current_device() == 0
        env.reset()
        policy = make_policy(15, seed=4, scale=1.5).to(dev)
        dpol = DevicePolicy.from_torch(policy.pi)
        assert dpol.device == torch.device(dev) and env.obs.device == torch.device(dev)
        outs = []
        for t in range(7):
            actions = dpol.act_env(env) if t % 2 else dpol.act(env.obs)
            obs, rew, done, info = env.step(actions)
            outs.append((obs.cpu(), rew.cpu(), done.cpu(), info.cpu(), actions.cpu()))
        assert torch.cuda.current_device() == 0
        results.append(outs)
        dpol.close()
        env.close()
    for a, b in zip(*results):
        for x, y in zip(a, b):
            assert torch.equal(x, y) or (torch.isnan(x) == torch.isnan(y)).all() and torch.equal(torch.nan_to_num(x), torch.nan_to_num(y))


def test_outputs_stay_inside_their_buffers():
    """Guard bands around caller-owned outputs (compute-sanitizer is not available on the GPU pool): the policy
    kernel's action vector (dense and ring front ends, ragged row counts) and the observation matrix of the three
    step paths (warp-per-env, resident-minibatch pipeline, tcgen05 pipeline) written through `obs_out`."""
    pad = 1024
    policy = make_policy(15, seed=6)
    dev = DevicePolicy.from_torch(policy.pi)
    for rows in (1, 127, 128, 129, 100003):
        obs = torch.randn(rows, 15, device='cuda')
        buf = torch.full((rows + 2 * pad,), float('nan'), device='cuda')
        dev.act(obs, buf[pad:pad + rows])
        assert torch.isnan(buf[:pad]).all() and torch.isnan(buf[pad + rows:]).all() and not torch.isnan(buf[pad:pad + rows]).any(), rows
    rng = np.random.RandomState(0)
    shapes = [(ProblemSpec('softmax', 4, (), 3), 150, 7), (ProblemSpec('softmax', 784, (), 10), 320, 5),
              (ProblemSpec('softmax', 784, (64,), 10), 320, 3)]
    for spec, rows, envs in shapes:
        feats = rng.uniform(size=(rows, spec.num_features)).astype(np.float32)
        labels = rng.randint(0, spec.num_outputs, rows).astype(np.int32)
        env = BatchedOptEnv(spec, feats, labels, envs, batch_size=32, max_batches=3, perms=env_permutations(rows, list(range(envs))))
        env.reset()
        n = env.num_rows * env.obs_dim
        buf = torch.full((n + 2 * pad,), float('nan'), device='cuda')
        window = buf[pad:pad + n].view(env.num_rows, env.obs_dim)
        gen = torch.Generator(device='cuda').manual_seed(1)
        for t in range(4):                                     # the third step ends every episode: auto-reset rows too
            env.step(torch.rand(env.num_rows, device='cuda', generator=gen) * 2, obs_out=window)
            assert torch.isnan(buf[:pad]).all() and torch.isnan(buf[pad + n:]).all() and not torch.isnan(window).any(), (spec, t)
        if spec.hidden:
            act = torch.full((env.num_rows + 2 * pad,), float('nan'), device='cuda')
            dev.act_env(env, act[pad:pad + env.num_rows])
            assert torch.isnan(act[:pad]).all() and torch.isnan(act[pad + env.num_rows:]).all()
            assert not torch.isnan(act[pad:pad + env.num_rows]).any()
        env.close()
    dev.close()


def test_graph_captured_rollout_equals_the_eager_rollout():
    """policy -> ring-only step -> bookkeeping captured as one CUDA graph of two steps and replayed: the same returns,
    episode count and parameters as the eager loop (deterministic policy); with exploration noise the device seed
    counter makes every replay draw fresh noise (the actions of consecutive steps differ)."""
    policy = make_policy(15, seed=2, scale=1.5)
    results = []
    for use_graph in (False, True):
        env = small_env(num_envs=4, max_batches=5, materialize_obs=False)
        dev = DevicePolicy.from_torch(policy.pi)
        returns, finished = device_policy_rollout(env, dev, 8, ring_only=True, use_graph=use_graph)
        results.append((returns.cpu().numpy(), finished, env.get_state('params').cpu().numpy()))
        dev.close()
        env.close()
    assert results[0][1] == results[1][1] == 4
    assert np.array_equal(results[0][0], results[1][0]) and np.array_equal(results[0][2], results[1][2])
    # noise: replays of the same graph must not repeat the draw
    env = small_env(num_envs=2, max_batches=50, materialize_obs=False)
    dev = DevicePolicy.from_torch(policy.pi, log_std=torch.tensor(-1.0))
    dev.use_seed_counter(True)
    actions = torch.empty(env.num_rows, device='cuda')
    graph = torch.cuda.CUDAGraph()
    dev.act_env(env, actions, seed=0)
    torch.cuda.synchronize()
    with torch.cuda.graph(graph):
        dev.act_env(env, actions, seed=0)
        dev.seed_counter.add_(1)
    graph.replay()
    first = actions.clone()
    graph.replay()
    assert not torch.equal(first, actions) and int(dev.seed_counter.item()) == 2
    dev.close()
    env.close()

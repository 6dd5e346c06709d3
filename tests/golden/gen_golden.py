"""Generate golden vectors by running the REFERENCE'S OWN env / history / vectoriser code.

Run in the build container only (needs /root/reference):

    python tests/golden/gen_golden.py

What executes from /root/reference, unmodified, imported by path:
  custom_envs.envs.multioptlrs.MultiOptLRs, custom_envs.envs.multioptimize.MultiOptimize,
  custom_envs.envs.baseenvironment, custom_envs.utils.utils_env, utils_common.History,
  utils_math.use_random_state, custom_envs.dataset.InMemoryDataSet,
  custom_envs.vectorize.optvecenv.OptVecEnv (thread-per-env pipes, auto-reset).
What is stubbed, because it is third-party and absent from the image:
  gym / stable_baselines (the stand-ins in custom_envs_b200/compat), numexpr,
  tensorflow.  The TensorFlow problem ``OptimizeNN`` is replaced by ``NumpyProblem``
  below: the reference's BaseProblem interface over the oracle's float64 loss/gradient
  (that arithmetic is pinned separately against torch.autograd in
  tests/test_oracle.py).  ``NumpyProblem.next/reset`` follow
  problems/optimize_nn.py:102-120 on the reference's real InMemoryDataSet, so the
  index stream (including the "same permutation every epoch" quirk of
  use_random_state) comes from reference code, not from the oracle.

The script writes tests/golden/*.npz; tests/test_oracle.py replays the oracle against
them.  Nothing under tests/ reads /root/reference at run time.
"""
import os
import sys
import threading
import types

import numpy as np
import numpy.random as npr

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REFERENCE = '/root/reference'


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()


class _Permissive(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _Anything


class _Sequence:
    """keras.utils.Sequence iteration protocol."""

    def __iter__(self):
        for item in (self[i] for i in range(len(self))):
            yield item


def install_stubs():
    sys.path.insert(0, REPO)
    from custom_envs_b200.compat import gym_standin, vec_env
    gym_standin.install_as_gym()
    for name in ('tensorflow', 'tensorflow.keras', 'tensorflow.keras.utils',
                 'tensorflow.keras.models', 'tensorflow.keras.layers', 'numexpr',
                 'stable_baselines', 'stable_baselines.common'):
        sys.modules[name] = _Permissive(name)
    sys.modules['tensorflow'].keras = sys.modules['tensorflow.keras']
    sys.modules['tensorflow.keras.utils'].Sequence = _Sequence
    sbv = types.ModuleType('stable_baselines.common.vec_env')
    sbv.VecEnv, sbv.CloudpickleWrapper = vec_env.VecEnv, vec_env.CloudpickleWrapper
    sbt = types.ModuleType('stable_baselines.common.tile_images')
    sbt.tile_images = vec_env.tile_images
    sys.modules['stable_baselines.common.vec_env'] = sbv
    sys.modules['stable_baselines.common.tile_images'] = sbt
    sys.path.insert(0, REFERENCE)          # 'custom_envs' now resolves to the reference


install_stubs()

import custom_envs                                             # noqa: E402 (reference)
from custom_envs.dataset import InMemoryDataSet                # noqa: E402
from custom_envs.problems.base_problem import BaseProblem      # noqa: E402
from custom_envs.envs import multioptlrs as ref_multioptlrs    # noqa: E402
from custom_envs.envs import multioptimize as ref_multioptimize  # noqa: E402
from custom_envs.vectorize.optvecenv import OptVecEnv          # noqa: E402
from oracle import optenv_oracle as orc                        # noqa: E402

assert custom_envs.__file__.startswith(REFERENCE)


class NumpyProblem(BaseProblem):
    """BaseProblem over the oracle's loss/grad; batching by the reference's data set."""

    def __init__(self, spec, feats, labels, batch_size, init_rng, log):
        self.spec = spec
        self.log = log
        self.init_rng = init_rng
        if spec.kind != 'func':
            ids = np.arange(len(feats), dtype=np.float64)[:, None]
            targets = np.concatenate([np.asarray(labels, np.float64).reshape(len(feats), -1),
                                      ids], axis=1)
            self.data_set = InMemoryDataSet(np.asarray(feats, np.float64), targets, batch_size)
        self.data_set_iter = None
        self.current_batch = None
        self.theta = None
        self.reset()                         # optimize_nn.py:64 (outside any env context)
        self.log['init_order'] = None if spec.kind == 'func' else \
            self.data_set.targets[:, -1].astype(np.int64).copy()
        self.log['resets'] = []              # constructor reset is not an env reset
        self.log['batches'] = []

    @property
    def size(self):
        return self.spec.size

    def next(self):                          # optimize_nn.py:102-112
        if self.spec.kind == 'func':
            return
        try:
            features, targets = next(self.data_set_iter)
        except StopIteration:
            self.data_set.on_epoch_end()
            self.data_set_iter = iter(self.data_set)
            features, targets = next(self.data_set_iter)
        self.current_batch = (features, targets)
        if 'batches' in self.log:
            self.log['batches'].append(targets[:, -1].astype(np.int64).copy())

    def reset(self):                         # optimize_nn.py:114-120
        self.theta = orc.glorot_uniform_init(self.spec, self.init_rng)
        if 'resets' in self.log:
            self.log['resets'].append(self.theta.copy())
        self.data_set_iter = iter(())
        self.next()

    def _eval(self):
        if self.spec.kind == 'func':
            grad, loss = orc.loss_and_grad(self.spec, self.theta[None], None, None, None)
        else:
            feats, targets = self.current_batch
            if self.spec.kind == 'softmax':
                targ = targets[:, 0].astype(np.int64)[None]
            else:
                targ = targets[:, :-1][None]
            mask = np.ones((1, len(feats)))
            grad, loss = orc.loss_and_grad(self.spec, self.theta[None],
                                           feats.astype(np.float32)[None], targ, mask)
        return grad[0].astype(np.float32), np.float32(loss[0])

    def get(self):                           # optimize_nn.py:152-159
        grad, loss = self._eval()
        return grad.astype(np.float64), loss, self.theta.astype(np.float64)

    def get_gradient(self):
        return self._eval()[0].astype(np.float64)

    def get_loss(self):
        return self._eval()[1]

    def get_parameters(self):
        return self.theta.astype(np.float64)

    def set_parameters(self, parameters):    # optimize_nn.py:142-150 (float32 feed)
        self.theta = np.asarray(parameters, np.float64).astype(np.float32)


_STEP_LOCK = threading.Lock()


def _locked(method):
    def call(*args, **kwargs):
        with _STEP_LOCK:
            return method(*args, **kwargs)
    return call


def env_perm(env, num_rows):
    """The permutation the env's (never advancing) RandomState yields under
    use_random_state (utils/utils_math.py:10-22)."""
    state = npr.RandomState()
    state.set_state(env.random_generator.get_state())
    idx = np.arange(num_rows)
    state.shuffle(idx)
    return idx


def run_case(name, env_kind, spec, num_rows, batch_size, num_envs, steps, env_kwargs,
             action_low, action_high, seed=0):
    data_rng = npr.RandomState(seed)
    if spec.kind == 'func':
        feats = labels = None
    else:
        feats = data_rng.uniform(size=(num_rows, spec.num_features))
        feats = ((feats - feats.min(0)) / (feats.max(0) - feats.min(0) + 1e-8)).astype(np.float32)
        if spec.kind == 'softmax':
            labels = np.arange(num_rows) % spec.num_outputs
        else:
            true_w = data_rng.normal(size=(spec.num_features, spec.num_outputs))
            labels = (feats @ true_w + 0.1 * data_rng.normal(size=(num_rows, spec.num_outputs))
                      ).astype(np.float32)
    logs = [dict() for _ in range(num_envs)]
    built = []

    def make_env(i):
        def factory():
            npr.seed(1000 + i)               # constructor shuffle uses the global RNG
            log = logs[i]
            problem = NumpyProblem(spec, feats, labels, batch_size,
                                   npr.RandomState(100 + i), log)
            if env_kind == 'optlrs':
                ref_multioptlrs.get_problem = lambda *a, **k: problem
                env = ref_multioptlrs.MultiOptLRs(problem='nn', **env_kwargs)
            else:
                ref_multioptimize.get_problem = lambda *a, **k: problem
                ref_multioptimize.load_data = lambda *a, **k: None
                kwargs = dict(env_kwargs)
                env = ref_multioptimize.MultiOptimize(**kwargs)
            env.seed(7 + i)
            # the reference's threads race on the global numpy RNG inside
            # use_random_state; serialise step/reset so the fixture is deterministic
            env.step = _locked(env.step)
            env.reset = _locked(env.reset)
            log['perm'] = None if spec.kind == 'func' else env_perm(env, num_rows)
            built.append(env)
            return env
        return factory

    # environments are built one after the other inside their worker threads; build them
    # serially here first so that the module-level get_problem patch is race free.
    envs = [make_env(i)() for i in range(num_envs)]
    vec = OptVecEnv([(lambda e=e: e) for e in envs])
    num_params = spec.size
    act_rng = npr.RandomState(2)
    out = {'states': [], 'rewards': [], 'dones': [], 'actions': []}
    info_keys = orc.INFO_KEYS
    infos_out = {k: [] for k in info_keys}
    infos_out['episode_r'], infos_out['episode_l'] = [], []
    out['reset_states'] = np.asarray(vec.reset(), np.float64)
    for _ in range(steps):
        actions = act_rng.uniform(action_low, action_high,
                                  size=(num_envs * num_params, 1)).astype(np.float32)
        states, rewards, dones, infos = vec.step(actions)
        out['actions'].append(actions[:, 0])
        out['states'].append(np.asarray(states, np.float64))
        out['rewards'].append(np.asarray(rewards, np.float64))
        out['dones'].append(np.asarray(dones, bool))
        for k in info_keys:
            vals = [infos[e * num_params][k] for e in range(num_envs)]
            infos_out[k].append([np.nan if v is None else float(v) for v in vals])
        infos_out['episode_r'].append(
            [float(infos[e * num_params]['episode']['r']) for e in range(num_envs)])
        infos_out['episode_l'].append(
            [int(infos[e * num_params]['episode']['l']) for e in range(num_envs)])
    vec.close()
    fixture = {
        'env_kind': env_kind, 'problem_kind': spec.kind,
        'num_features': spec.num_features, 'hidden': np.asarray(spec.hidden, np.int64),
        'num_outputs': spec.num_outputs, 'num_rows': num_rows,
        'batch_size': -1 if batch_size is None else batch_size, 'num_envs': num_envs,
        'env_kwargs_keys': np.asarray(list(env_kwargs.keys())),
        'env_kwargs_vals': np.asarray(list(env_kwargs.values()), np.int64),
        'actions': np.asarray(out['actions'], np.float32),
        'reset_states': out['reset_states'],
        'states': np.asarray(out['states']), 'rewards': np.asarray(out['rewards']),
        'dones': np.asarray(out['dones']),
    }
    for k, v in infos_out.items():
        fixture['info_' + k] = np.asarray(v, np.float64)
    if spec.kind != 'func':
        fixture['feats'] = feats
        fixture['labels'] = np.asarray(labels)
        fixture['perms'] = np.stack([log['perm'] for log in logs])
        fixture['init_orders'] = np.stack([log['init_order'] for log in logs])
        nb = max(len(log['batches']) for log in logs)
        bsz = num_rows if batch_size is None else batch_size
        batches = np.full((num_envs, nb, bsz), -1, np.int64)
        for e, log in enumerate(logs):
            for b, ids in enumerate(log['batches']):
                batches[e, b, :len(ids)] = ids
        fixture['batches'] = batches
    nres = max(len(log['resets']) for log in logs)
    resets = np.full((num_envs, nres, num_params), np.nan, np.float32)
    for e, log in enumerate(logs):
        for r, theta in enumerate(log['resets']):
            resets[e, r] = theta
    fixture['reset_params'] = resets
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **fixture)
    print('wrote', path, {k: getattr(v, 'shape', v) for k, v in fixture.items()
                          if k in ('states', 'batches', 'reset_params')})


def main():
    lrs = dict(max_batches=12, max_history=5)
    run_case('optlrs_softmax_iris', 'optlrs',
             orc.ProblemSpec('softmax', 4, (), 3), 150, 32, 2, 30, lrs, 0.0, 3.5)
    run_case('optlrs_mlp_small', 'optlrs',
             orc.ProblemSpec('softmax', 6, (5,), 3), 50, 16, 2, 30,
             dict(max_batches=9, max_history=3), 0.0, 3.0)
    run_case('optlrs_linreg', 'optlrs',
             orc.ProblemSpec('linreg', 4, (), 1), 150, 32, 1, 100,
             dict(max_batches=100, max_history=5), 0.0, 2.5)
    run_case('optlrs_func', 'optlrs',
             orc.ProblemSpec('func', 0, (), 0), 0, None, 2, 25,
             dict(max_batches=10, max_history=5), -1.0, 1.5)
    run_case('optlrs_diverge', 'optlrs',
             orc.ProblemSpec('softmax', 4, (), 3), 150, 32, 2, 20,
             dict(max_batches=15, max_history=5), 4.5, 6.0)
    for hv in range(5):
        for ov in range(4):
            run_case('optimize_h%d_o%d' % (hv, ov), 'optimize',
                     orc.ProblemSpec('softmax', 4, (), 3), 150, 32, 2, 14,
                     dict(version=hv, observation_version=ov, action_version=hv % 2,
                          reward_version=(hv + ov) % 7, max_batches=6, max_history=4),
                     -1.0, 1.0)


if __name__ == '__main__':
    main()

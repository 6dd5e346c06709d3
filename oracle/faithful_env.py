"""Reference-faithful CPU path of the optimise-env step (TEST / BASELINE INFRASTRUCTURE, never on the
product path): one Python env OBJECT per env, driven one thread per env over pipes, the way the
reference runs (SURVEY 8d path (i)).

What is restated here and from where:
* ``CpuProblem``     -- ``OptimizeNN``'s BaseProblem surface (problems/optimize_nn.py:102-159) over the
                        oracle's numpy float32 loss / batch-SUM gradient (TensorFlow is not installable)
                        and ``InMemoryDataSet``'s slices + same-permutation reshuffle
                        (dataset/inmemorydataset.py:11-28, utils/utils_math.py:10-22);
* ``CpuMultiOptLRs`` -- ``MultiOptLRs.base_reset / base_step`` (envs/multioptlrs.py:39-129) and
                        ``BaseEnvironment.step / reset`` (envs/baseenvironment.py:30-49): the per-agent
                        action dict in, the per-agent observation dict out (one ``np.clip(np.nan_to_num(
                        list(row)))`` PER AGENT, which is where the reference spends its time), the raw and
                        the adjusted ``History``, reward, divergence rule, the 14 info statistics.
``History``, ``get_observation`` / ``get_reward`` / ``get_action_optlrs`` and the thread-per-env pipe
vectoriser (``ThreadVecEnv`` + ``OptEnvRunner`` behind ``OptVecEnv``'s generic path) are the host
mirrors under custom_envs_b200/, themselves checked against the reference's own modules imported by
path (tests/test_oracle.py, tests/test_host_layer.py).  ``/root/reference`` is never read at run time.
"""
import numpy as np

from custom_envs_b200.compat import spaces
from custom_envs_b200.compat.gym_standin import Env, np_random
from custom_envs_b200.utils import utils_env
from custom_envs_b200.utils.utils_common import History
from oracle import optenv_oracle as orc

BOUNDS = 100
AGENT_FMT = 'parameter-{:d}'


class CpuProblem:
    """One model + its private minibatch stream; flat float64 vectors in and out like
    ``flatten_arrays`` (utils/utils_common.py:199-207), float32 arithmetic inside."""

    def __init__(self, spec, feats, targs, batch_size=32, seed=0, compute_dtype=np.float32):
        self.spec, self.dtype = spec, compute_dtype
        self.feats, self.targs = np.asarray(feats, np.float32), np.asarray(targs)
        self.stream = orc.IndexStream(len(self.feats), batch_size, orc.env_permutation(len(self.feats), seed)[None])
        self.rng = np.random.RandomState(seed)
        self.size = spec.size
        self.parameters = orc.glorot_uniform_init(spec, self.rng).astype(np.float64)
        self._everyone = np.ones(1, bool)

    def reset(self):
        self.parameters = orc.glorot_uniform_init(self.spec, self.rng).astype(np.float64)
        self.stream.reset(self._everyone)

    def next(self):
        self.stream.advance(self._everyone)

    def get(self):
        idx, cnt = self.stream.current()
        mask = np.arange(idx.shape[1])[None, :] < cnt[:, None]
        grad, loss = orc.loss_and_grad(self.spec, self.parameters[None].astype(np.float32), self.feats[idx],
                                       self.targs[idx], mask, self.dtype)
        return (grad[0].astype(np.float32).astype(np.float64), np.float32(loss[0]), self.parameters.copy())

    def get_gradient(self):
        return self.get()[0]

    def get_loss(self):
        return self.get()[1]

    def get_parameters(self):
        return self.parameters.copy()

    def set_parameters(self, parameters):
        self.parameters = np.asarray(parameters, np.float64).astype(np.float32).astype(np.float64)


class CpuMultiOptLRs(Env):
    """The multi-agent learning-rate env as a per-env Python object."""

    def __init__(self, problem, max_batches=400, max_history=5, seed=0):
        self.model = problem
        size = problem.size
        self.history = History(5, losses=(), gradients=(size,), weights=(size,))
        obs_space, self.adjusted_history = utils_env.get_obs_version((size,), max_history, 3)
        act_space = utils_env.get_action_space_optlrs(2)
        self._names = [AGENT_FMT.format(i) for i in range(size)]
        self.observation_space = spaces.Dict({name: obs_space for name in self._names})
        self.action_space = spaces.Dict({name: act_space for name in self._names})
        self.max_batches, self.max_history = max_batches, max_history
        self.current_step = 0
        self.random_generator, _ = np_random(seed)

    def _agent_rows(self, rows):
        out = {}
        for name, row in zip(self._names, rows):                  # one small numpy round trip per agent
            out[name] = np.clip(np.nan_to_num(list(row)), -BOUNDS, BOUNDS) - 1
        return out

    def reset(self):
        self.current_step = 0
        self.adjusted_history.reset()
        self.model.reset()
        self.history.reset()
        grad, loss, weights = self.model.get()
        self.history.append(losses=loss, gradients=grad, weights=weights)
        return self._agent_rows(self.adjusted_history.build_multistate())

    def step(self, action):
        self.current_step += 1
        flat = np.reshape([np.ravel(action[name]) for name in self._names], (-1,))
        grad0 = self.model.get_gradient()
        rates = utils_env.get_action_optlrs(flat, 0)
        self.model.set_parameters(self.model.parameters - grad0 * rates)
        grad, loss, weights = self.model.get()
        self.history.append(losses=loss, gradients=grad, weights=weights)
        adj_loss, adj_wght, adj_grad = utils_env.get_observation(self.history, 3)
        self.adjusted_history.append(weights=adj_wght, losses=adj_loss, gradients=adj_grad)
        state = self.adjusted_history.build_multistate()
        states = self._agent_rows(state)
        reward = np.clip(utils_env.get_reward(loss, adj_loss, 6), -BOUNDS, BOUNDS)
        terminal = self.current_step >= self.max_batches
        if not terminal and loss > 1e4:
            terminal = True
            reward -= self.max_batches - self.current_step
        past = self.history['gradients']
        info = {
            'loss': self.model.get_loss() if terminal else None,
            'batch_loss': loss,
            'weights_mean': np.mean(np.abs(weights)), 'weights_sum': np.sum(np.abs(weights)),
            'actions_mean': np.mean(rates), 'actions_std': np.std(rates),
            'states_mean': np.mean(np.abs(state)), 'states_sum': np.sum(np.abs(state)),
            'grads_mean': np.mean(past), 'grads_sum': np.sum(past),
            'loss_mean': np.mean(self.history['losses']),
            'adjusted_loss': float(adj_loss), 'adjusted_grad': np.mean(np.abs(adj_grad)),
            'grad_diff': np.mean(np.abs(past[0] - past[1])),
            'episode': {'r': reward, 'l': self.current_step},
        }
        self.model.next()
        return states, reward, terminal, info


def make_vec_env(spec, feats, targs, num_envs, batch_size=32, max_batches=400, max_history=5):
    """``OptVecEnv`` over ``num_envs`` per-env objects: thread per env, pipes, rows = agents."""
    from custom_envs_b200.vectorize.optvecenv import OptVecEnv

    def factory(i):
        return lambda: CpuMultiOptLRs(CpuProblem(spec, feats, targs, batch_size, seed=i), max_batches,
                                      max_history, seed=i)
    return OptVecEnv([factory(i) for i in range(num_envs)])


def env_steps_per_s(spec, feats, targs, num_envs, steps, warmup=1, **kwargs):
    """Timed loop of the faithful path: (env-steps/s, seconds)."""
    import time
    vec = make_vec_env(spec, feats, targs, num_envs, **kwargs)
    vec.reset()
    actions = np.random.RandomState(2).uniform(0, 3, size=(vec.num_envs, 1)).astype(np.float32)
    for _ in range(warmup):
        vec.step(actions)
    t0 = time.perf_counter()
    for _ in range(steps):
        vec.step(actions)
    elapsed = time.perf_counter() - t0
    vec.close()
    return num_envs * steps / elapsed, elapsed

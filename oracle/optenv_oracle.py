"""CPU oracle for the batched optimise-env step.  TEST INFRASTRUCTURE ONLY.

This module restates, in numpy, the algorithm of the reference hot path
(adolfogonzalez3/custom_envs).  It is the *checker* for the CUDA product path in
``custom_envs_b200``; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product path never routes through this file.

Parity status
-------------
* History layout / observation ratios / reward / action transforms / row order /
  RNG-context semantics are PINNED: ``tests/golden/gen_golden.py`` drives the
  reference's own ``utils_common.History``, ``utils_env`` and
  ``utils_math.use_random_state`` (imported by path from /root/reference) next to
  this restatement and commits the outputs under ``tests/golden/``.
* The forward/backward arithmetic lives in TensorFlow 1.x (not vendored, not
  installable here, version unpinned by the reference: install_conda.sh:8).  It is
  restated from the published definitions of ``keras.layers.Dense`` (Glorot-uniform
  kernel, zero bias), ``relu``, ``softmax`` + ``categorical_crossentropy`` and
  ``tf.gradients`` of a per-sample loss vector (= batch SUM) and cross-checked against
  ``torch.autograd`` in float64.  Against TensorFlow itself: PARITY UNPINNED.

Reference call sites restated (all paths relative to /root/reference):
  problems/optimize_nn.py:35-64,102-120,152-159   graph, next/reset, get
  utils/utils_tf.py:74-86                          Dense(h, relu) stack
  utils/utils_common.py:12-23,102-196,199-225      shuffle, History, flatten
  utils/utils_math.py:10-22                        use_random_state
  utils/utils_env.py:9-164                         version tables
  dataset/inmemorydataset.py:11-28                 batch slicing
  envs/baseenvironment.py:30-49                    step counter, episode info
  envs/multioptlrs.py:39-129                       MultiOptLRs
  envs/multioptimize.py:40-154                     MultiOptimize
  vectorize/optvecenv.py:10-91, concurrentvecenv.py:32-38   row order, auto-reset
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

BOUNDS = 1e2            # envs/multioptlrs.py:16
RAW_DEPTH = 5           # envs/multioptlrs.py:42 (raw History depth is fixed at 5)
INFO_KEYS = (           # envs/multioptlrs.py:111-127, in insertion order
    'loss', 'batch_loss', 'weights_mean', 'weights_sum', 'actions_mean',
    'actions_std', 'states_mean', 'states_sum', 'grads_mean', 'grads_sum',
    'loss_mean', 'adjusted_loss', 'adjusted_grad', 'grad_diff')


# --------------------------------------------------------------------------- rows
def lexicographic_rows(num_params: int) -> np.ndarray:
    """perm[j] = parameter index shown in VecEnv row j.

    gym's ``spaces.Dict`` sorts its keys and ``flatten_dictionary`` sorts the
    observation dict by full name (vectorize/optvecenv.py:10-14), so agents appear in
    lexicographic order of ``'parameter-%d'`` (envs/baseenvironment.py:64).
    """
    names = np.array(['parameter-%d' % i for i in range(num_params)])
    return np.argsort(names, kind='stable').astype(np.int64)


# ------------------------------------------------------------------------ problems
@dataclass(frozen=True)
class ProblemSpec:
    """Shape of one optimisation problem.

    kind: 'softmax' (Dense stack + softmax/CE, problems/optimize_nn.py:35-52),
          'linreg'  (XW+b with utils_math.mse per-sample loss, utils/utils_math.py:37-48;
                     not a reference problem),
          'func'    (Rosenbrock, problems/optimize_function.py:35-50).
    """
    kind: str = 'softmax'
    num_features: int = 4
    hidden: tuple = ()
    num_outputs: int = 3

    @property
    def shapes(self):
        """Variable shapes in creation order: kernel[in,out], bias[out] per layer."""
        if self.kind == 'func':
            return ((), ())
        dims = (self.num_features,) + tuple(self.hidden) + (self.num_outputs,)
        out = []
        for a, b in zip(dims[:-1], dims[1:]):
            out += [(a, b), (b,)]
        return tuple(out)

    @property
    def size(self):
        if self.kind == 'func':
            return 2
        return int(sum(int(np.prod(s)) for s in self.shapes))


def glorot_uniform_init(spec: ProblemSpec, rng: np.random.RandomState) -> np.ndarray:
    """keras Dense defaults: Glorot-uniform kernels, zero biases; flat float32."""
    if spec.kind == 'func':
        return np.array([-1.9, 2.0], np.float32)       # optimize_function.py:37
    parts = []
    for shape in spec.shapes:
        if len(shape) == 2:
            limit = np.sqrt(6.0 / (shape[0] + shape[1]))
            parts.append(rng.uniform(-limit, limit, size=shape).ravel())
        else:
            parts.append(np.zeros(shape))
    return np.concatenate(parts).astype(np.float32)


def _unflatten(spec, theta):
    """theta: [E, P] -> list of [E, *shape] (utils/utils_common.py:210-225)."""
    arrays, start = [], 0
    for shape in spec.shapes:
        n = int(np.prod(shape))
        arrays.append(theta[:, start:start + n].reshape((theta.shape[0],) + shape))
        start += n
    return arrays


def loss_and_grad(spec: ProblemSpec, theta, feats, targs, mask, dtype=np.float64, abs_terms=False):
    """Batched forward/backward.

    theta [E,P]; feats [E,B,D]; targs [E,B] int labels ('softmax') or [E,B,C] floats
    ('linreg'); mask [E,B] 1.0 for valid rows (ragged last batch,
    dataset/inmemorydataset.py:24-28).
    Returns (grad [E,P] = d(SUM_i loss_i)/d theta, loss [E] = MEAN_i loss_i), which is
    what ``tf.gradients(loss_vec, w)`` and ``tf.reduce_mean(loss_vec)`` produce
    (problems/optimize_nn.py:47-52).
    ``abs_terms=True`` appends S [E,P], the sum of the ABSOLUTE terms of the last dot product behind
    every gradient element (|activations|^T . |deltas|): the scale fp32 rounding noise of that
    element is proportional to, which the parity tests bound the device's error with.
    """
    theta = np.asarray(theta, dtype)
    num_envs = theta.shape[0]
    if spec.kind == 'func':
        x, y = theta[:, 0], theta[:, 1]
        loss = 100 * (y - x ** 2) ** 2 + (1 - x) ** 2       # utils_functions.py:4-6
        grad = np.stack([-400 * x * (y - x ** 2) - 2 * (1 - x), 200 * (y - x ** 2)], 1)
        return grad, loss
    feats = np.asarray(feats, dtype)
    mask = np.asarray(mask, dtype)
    count = mask.sum(axis=1)
    params = _unflatten(spec, theta)
    acts, pres = [feats], []
    cur = feats
    nlayers = len(params) // 2
    for li in range(nlayers):
        kern, bias = params[2 * li], params[2 * li + 1]
        cur = np.matmul(cur, kern) + bias[:, None, :]
        pres.append(cur)
        if li < nlayers - 1:
            cur = np.maximum(cur, 0)                         # relu, utils_tf.py:74
            acts.append(cur)
    out = cur
    if spec.kind == 'softmax':
        zmax = out.max(axis=2, keepdims=True)
        ez = np.exp(out - zmax)
        sez = ez.sum(axis=2, keepdims=True)
        logp = out - zmax - np.log(sez)
        labels = np.asarray(targs, np.int64)
        onehot = np.zeros_like(out)
        np.put_along_axis(onehot, labels[..., None], 1.0, axis=2)
        per_sample = -(onehot * logp).sum(axis=2)
        dout = (ez / sez - onehot)
    elif spec.kind == 'linreg':
        diff = out - np.asarray(targs, dtype)
        per_sample = 0.5 * (diff ** 2).sum(axis=2)
        dout = diff
    else:
        raise RuntimeError('Not a name of a problem.')
    loss = (per_sample * mask).sum(axis=1) / count
    dout = dout * mask[..., None]
    grads = [None] * len(params)
    sums = [None] * len(params)
    for li in reversed(range(nlayers)):
        grads[2 * li] = np.matmul(acts[li].transpose(0, 2, 1), dout)
        grads[2 * li + 1] = dout.sum(axis=1)
        if abs_terms:
            sums[2 * li] = np.matmul(np.abs(acts[li]).transpose(0, 2, 1), np.abs(dout))
            sums[2 * li + 1] = np.abs(dout).sum(axis=1)
        if li > 0:
            dout = np.matmul(dout, params[2 * li].transpose(0, 2, 1))
            dout = dout * (pres[li - 1] > 0)
    grad = np.concatenate([g.reshape(num_envs, -1) for g in grads], axis=1)
    if abs_terms:
        return grad, loss, np.concatenate([t.reshape(num_envs, -1) for t in sums], axis=1)
    return grad, loss


# -------------------------------------------------------------------- index stream
def env_permutation(num_rows: int, seed) -> np.ndarray:
    """The permutation every epoch-end shuffle of one env applies.

    ``use_random_state`` (utils/utils_math.py:10-22) loads a COPY of the env's
    RandomState into the global generator, so the env's generator never advances and
    every ``shuffle`` (utils/utils_common.py:12-23) inside ``step``/``reset``
    (envs/baseenvironment.py:38,48) draws the same permutation.
    """
    indices = np.arange(num_rows)
    np.random.RandomState(seed).shuffle(indices)
    return indices


class IndexStream:
    """Minibatch row indices of E envs (dataset/inmemorydataset.py, optimize_nn.py:102-120)."""

    def __init__(self, num_rows, batch_size, perms, init_orders=None):
        self.num_rows = int(num_rows)
        self.batch_size = self.num_rows if batch_size is None else int(batch_size)
        self.perms = np.asarray(perms, np.int64).reshape(-1, self.num_rows)
        num_envs = self.perms.shape[0]
        if init_orders is None:
            init_orders = np.tile(np.arange(self.num_rows), (num_envs, 1))
        self.orders = np.array(init_orders, np.int64).reshape(num_envs, self.num_rows)
        self.cursor = np.zeros(num_envs, np.int64)
        self.num_batches = -(-self.num_rows // self.batch_size)

    def _shuffle(self, env_mask):
        for e in np.nonzero(env_mask)[0]:
            self.orders[e] = self.orders[e][self.perms[e]]

    def reset(self, env_mask):
        """``OptimizeNN.reset`` starts from an exhausted iterator -> reshuffle."""
        self._shuffle(env_mask)
        self.cursor[env_mask] = 0

    def advance(self, env_mask):
        """``OptimizeNN.next``: next slice, reshuffle + restart when exhausted."""
        self.cursor[env_mask] += 1
        wrapped = env_mask & (self.cursor >= self.num_batches)
        self._shuffle(wrapped)
        self.cursor[wrapped] = 0

    def current(self):
        """-> (idx [E,B] int32, padded with 0; cnt [E] int32)."""
        num_envs = self.orders.shape[0]
        idx = np.zeros((num_envs, self.batch_size), np.int32)
        cnt = np.zeros(num_envs, np.int32)
        for e in range(num_envs):
            lo = self.cursor[e] * self.batch_size
            rows = self.orders[e][lo:lo + self.batch_size]
            idx[e, :len(rows)] = rows
            cnt[e] = len(rows)
        return idx, cnt


# ------------------------------------------------------------- version tables (E6)
def action_transform(action, version):
    """utils/utils_env.py:102-123 (MultiOptLRs family)."""
    if version == 0:
        return 10 ** (action - 4)
    if version == 1:
        return action * 1e-3
    if version == 2:
        return 2 ** action
    if version == 3:
        return np.clip((action + 1e3) * 1e-6, 0, np.inf)
    raise RuntimeError()


def delta_transform(action, version):
    """envs/multioptimize.py:95-102 (MultiOptimize)."""
    if version == 0:
        return np.sign(action) * 10 ** (np.abs(action) - 3)
    if version == 1:
        return action * 1e-3
    raise RuntimeError()


def reward_fn(loss, adj_loss, version):
    """utils/utils_env.py:71-99; vectorised over envs.  ``loss`` is TensorFlow's float32
    scalar in the reference, so ``1 / loss`` and ``np.log(1 / loss)`` are float32 there."""
    loss32 = np.asarray(loss, np.float32)
    with np.errstate(all='ignore'):
        if version == 0:
            return -adj_loss
        if version == 1:
            return (1 / loss32).astype(np.float64)
        if version == 2:
            return -adj_loss * 100
        if version == 3:
            return (1 / loss32).astype(np.float64) * 100
        if version == 4:
            return np.log(1 / loss32).astype(np.float64)
        if version == 5:
            return -(adj_loss - 1) ** 2
        if version == 6:
            return -(adj_loss - 1)
    raise RuntimeError()


def _loss_ratio(losses, version):
    with np.errstate(all='ignore'):
        if version in (0, 1):
            return losses[0] / (np.abs(losses[1]) + 1e-3)
        if version == 2:
            return (losses[0] - losses[1]) / (np.abs(losses[1] - losses[2]) + 1e-3)
        if version == 3:
            return np.nan_to_num(losses[0] / np.abs(losses[1]))
    raise RuntimeError()


def observation_ratios(losses, grads, weights, version, loss_is_f32=None):
    """utils/utils_env.py:126-164.  Inputs are raw histories, newest first:
    losses [R,E], grads/weights [R,E,P].  Returns (adj_loss [E], adj_w, adj_g [E,P]).

    ``loss_is_f32`` [E] marks envs whose raw loss History holds only float32 entries
    (every slot overwritten since the last ``History.reset``): ``History.__getitem__``
    (utils/utils_common.py:129-137) then yields a float32 array and the reference's
    loss ratio is evaluated in float32; until then the zero-filled float64 slots
    promote it to float64.  Weights/gradients are always float64
    (``flatten_arrays``, utils/utils_common.py:199-207)."""
    adj_l = _loss_ratio(losses, version)
    if loss_is_f32 is not None and np.any(loss_is_f32):
        adj_l32 = _loss_ratio(losses.astype(np.float32), version).astype(np.float64)
        adj_l = np.where(loss_is_f32, adj_l32, adj_l)
    with np.errstate(all='ignore'):
        adj_w = weights[0] / (np.abs(weights[1]) + 1e-3)
        adj_g = grads[0] / (np.abs(grads[1]) + 1e-3)
        if version == 0:
            pass
        elif version == 1:
            adj_g = grads[0] * 1e2
        elif version == 2:
            adj_w = (np.abs(weights[1] - weights[2])
                     / (np.abs(weights[0] - weights[1]) + 1e-8))
            adj_g = (grads[0] - grads[1]) / (np.abs(grads[1] - grads[2]) + 1e-3)
        elif version == 3:
            adj_g = np.nan_to_num(grads[0] / np.abs(grads[1]))
            adj_w = np.nan_to_num(weights[0] / np.abs(weights[1]))
        else:
            raise RuntimeError()
    return adj_l, adj_w, adj_g


# history layouts: utils/utils_env.py:22-44 -> (depth, keys in insertion order)
def history_layout(version, max_history):
    table = {
        0: (1, ('gradients',)),
        1: (max_history, ('losses', 'gradients')),
        2: (1, ('weights', 'losses', 'gradients')),
        3: (max_history, ('weights', 'losses', 'gradients')),
        4: (max_history, ('gradients',)),
    }
    if version not in table:
        raise RuntimeError()
    return table[version]


# ------------------------------------------------------------------------ the envs
@dataclass
class EnvConfig:
    env: str = 'optlrs'            # 'optlrs' (MultiOptLRs) | 'optimize' (MultiOptimize)
    max_batches: int = 400
    max_history: int = 5
    history_version: int = 3
    observation_version: int = 3
    action_version: int = 0
    reward_version: int = 6

    @classmethod
    def multioptlrs(cls, max_batches=400, max_history=5):
        """envs/multioptlrs.py:61 -> VersionType(3, 3, 0, 6)."""
        return cls('optlrs', max_batches, max_history, 3, 3, 0, 6)

    @classmethod
    def multioptimize(cls, version=1, max_batches=400, max_history=5,
                      observation_version=0, action_version=0, reward_version=0):
        """envs/multioptimize.py:40-42,73-75."""
        return cls('optimize', max_batches, max_history, version,
                   observation_version, action_version, reward_version)


class BatchedOptEnvOracle:
    """E independent optimise envs advanced in lock step, numpy, natural param order.

    ``compute_dtype`` float64 gives the parity oracle (loss/grad are then rounded to
    float32 like TensorFlow's outputs before the float64 env arithmetic of
    utils_env / History); float32 is used for the timed CPU baseline.
    """

    def __init__(self, spec: ProblemSpec, feats, targs, num_envs, batch_size=32,
                 config: EnvConfig = None, perms=None, seeds=None, init_orders=None,
                 compute_dtype=np.float64, init_seed=0):
        self.spec = spec
        self.cfg = config or EnvConfig.multioptlrs()
        self.num_envs = int(num_envs)
        self.compute_dtype = compute_dtype
        self.init_rng = np.random.RandomState(init_seed)
        if spec.kind != 'func':
            self.feats = np.asarray(feats, np.float32)
            self.targs = np.asarray(targs)
            num_rows = self.feats.shape[0]
            if perms is None:
                seeds = range(self.num_envs) if seeds is None else seeds
                perms = np.stack([env_permutation(num_rows, s) for s in seeds])
            self.stream = IndexStream(num_rows, batch_size, perms, init_orders)
        else:
            self.feats = self.targs = self.stream = None
        num_params = spec.size
        self.num_params = num_params
        depth, keys = history_layout(self.cfg.history_version, self.cfg.max_history)
        self.depth, self.keys = depth, keys
        self.obs_dim = depth * len(keys)
        e = self.num_envs
        self.weights = np.zeros((e, num_params), np.float32)
        self.raw_l = np.zeros((RAW_DEPTH, e))
        self.raw_g = np.zeros((RAW_DEPTH, e, num_params))
        self.raw_w = np.zeros((RAW_DEPTH, e, num_params))
        self.adj_l = np.zeros((depth, e))
        self.adj_g = np.zeros((depth, e, num_params))
        self.adj_w = np.zeros((depth, e, num_params))
        self.current_step = np.zeros(e, np.int64)
        self.raw_pushes = np.zeros(e, np.int64)     # appends since the raw History reset
        self.ext_idx = None

    # -- problem plumbing ------------------------------------------------------
    def set_batch(self, idx, cnt):
        """Pin the current minibatch (host-supplied indices, north_star protocol)."""
        self.ext_idx = (np.asarray(idx, np.int32), np.asarray(cnt, np.int32))

    def current_batch(self):
        if self.ext_idx is not None:
            return self.ext_idx
        return self.stream.current()

    def evaluate(self, env_mask=None):
        """``OptimizeNN.get``: (grad, loss) on the current batch, float32-rounded."""
        if self.spec.kind == 'func':
            grad, loss = loss_and_grad(self.spec, self.weights, None, None, None,
                                       self.compute_dtype)
        else:
            idx, cnt = self.current_batch()
            feats = self.feats[idx]
            targs = self.targs[idx]
            mask = (np.arange(idx.shape[1])[None, :] < cnt[:, None])
            grad, loss = loss_and_grad(self.spec, self.weights, feats, targs, mask,
                                       self.compute_dtype)
        return (grad.astype(np.float32).astype(np.float64),
                loss.astype(np.float32).astype(np.float64))

    # -- history helpers -------------------------------------------------------
    @staticmethod
    def _push(ring, value, env_mask):
        ring[1:, env_mask] = ring[:-1, env_mask]
        ring[0, env_mask] = value[env_mask]

    def _observation(self):
        """``History.build_multistate`` (utils/utils_common.py:188-196): per agent, for
        each key in insertion order, the ``depth`` newest-first values."""
        cols = []
        for key in self.keys:
            if key == 'weights':
                cols.append(self.adj_w.transpose(1, 2, 0))
            elif key == 'gradients':
                cols.append(self.adj_g.transpose(1, 2, 0))
            else:
                cols.append(np.broadcast_to(
                    self.adj_l.T[:, None, :],
                    (self.num_envs, self.num_params, self.depth)))
        state = np.concatenate(cols, axis=2)                 # [E,P,obs_dim]
        if self.cfg.env == 'optlrs':                         # multioptlrs.py:97-101
            obs = np.clip(np.nan_to_num(state), -BOUNDS, BOUNDS) - 1
        else:                                                # multioptimize.py:126-129
            obs = state
        return state, obs

    # -- gym surface -------------------------------------------------------------
    def reset(self, env_mask=None, init_params=None):
        """``base_reset`` (multioptlrs.py:66-78 / multioptimize.py:78-88) for the masked
        envs.  Returns obs [E,P,obs_dim] (rows of unmasked envs are their current obs)."""
        if env_mask is None:
            env_mask = np.ones(self.num_envs, bool)
        env_mask = np.asarray(env_mask, bool)
        self.current_step[env_mask] = 0                       # baseenvironment.py:47
        self.adj_l[:, env_mask] = 0
        self.adj_g[:, env_mask] = 0
        self.adj_w[:, env_mask] = 0
        for e in np.nonzero(env_mask)[0]:
            if init_params is not None:
                self.weights[e] = init_params[e]
            else:
                self.weights[e] = glorot_uniform_init(self.spec, self.init_rng)
        if self.stream is not None and self.ext_idx is None:
            self.stream.reset(env_mask)
        if self.cfg.env == 'optlrs':                          # multioptlrs.py:69
            self.raw_l[:, env_mask] = 0
            self.raw_g[:, env_mask] = 0
            self.raw_w[:, env_mask] = 0
            self.raw_pushes[env_mask] = 0
        self.raw_pushes[env_mask] += 1
        grad, loss = self.evaluate()
        self._push(self.raw_l, loss, env_mask)
        self._push(self.raw_g, grad, env_mask)
        self._push(self.raw_w, self.weights.astype(np.float64), env_mask)
        return self._observation()[1]

    def step(self, actions, advance=True):
        """``base_step`` for all envs.  actions [E,P] in NATURAL parameter order.
        Returns obs [E,P,obs_dim], reward [E], done [E] bool, info dict of [E] arrays
        (info['loss'] is NaN where the reference has None)."""
        cfg = self.cfg
        everyone = np.ones(self.num_envs, bool)
        self.current_step += 1                                # baseenvironment.py:37
        actions = np.asarray(actions).reshape(self.num_envs, self.num_params)
        w_prev = self.weights.astype(np.float64)
        with np.errstate(all='ignore'):
            if cfg.env == 'optlrs':
                grad0, _ = self.evaluate()                    # multioptlrs.py:85
                act = action_transform(actions, cfg.action_version)
                new_w = w_prev - grad0 * act                  # :87
            else:
                act = delta_transform(actions, cfg.action_version)
                new_w = w_prev - act                          # multioptimize.py:103
            self.weights = new_w.astype(np.float32)          # float32 placeholders
            grad, loss = self.evaluate()                      # :88
            weights = self.weights.astype(np.float64)
            self._push(self.raw_l, loss, everyone)
            self._push(self.raw_g, grad, everyone)
            self._push(self.raw_w, weights, everyone)
            self.raw_pushes += 1
            loss_is_f32 = self.raw_pushes >= RAW_DEPTH
            adj_l, adj_w, adj_g = observation_ratios(
                self.raw_l, self.raw_g, self.raw_w, cfg.observation_version, loss_is_f32)
            if 'losses' in self.keys:
                self._push(self.adj_l, adj_l, everyone)
            if 'weights' in self.keys:
                self._push(self.adj_w, adj_w, everyone)
            if 'gradients' in self.keys:
                self._push(self.adj_g, adj_g, everyone)
            state, obs = self._observation()
            reward = reward_fn(loss, adj_l, cfg.reward_version)
            done = self.current_step >= cfg.max_batches
            if cfg.env == 'optlrs':
                reward = np.clip(reward, -BOUNDS, BOUNDS)     # multioptlrs.py:103
                diverged = (~done) & (loss > 1e4)             # :105-107
                reward = np.where(diverged,
                                  reward - (cfg.max_batches - self.current_step), reward)
                done = done | diverged
            # the reference reduces the transformed action in its own dtype (float32
            # when the policy hands over float32 actions)
            act_rows = [np.asarray(row) for row in act]
            info = {
                'loss': np.where(done, loss, np.nan),         # get_loss(): same batch, w_t
                'batch_loss': loss,
                'weights_mean': np.mean(np.abs(weights), axis=1),
                'weights_sum': np.sum(np.abs(weights), axis=1),
                'actions_mean': np.array([np.mean(row) for row in act_rows], np.float64),
                'actions_std': np.array([np.std(row) for row in act_rows], np.float64),
                'states_mean': np.mean(np.abs(state), axis=(1, 2)),
                'states_sum': np.sum(np.abs(state), axis=(1, 2)),
                'grads_mean': np.mean(self.raw_g, axis=(0, 2)),
                'grads_sum': np.sum(self.raw_g, axis=(0, 2)),
                'loss_mean': np.where(
                    loss_is_f32,
                    np.mean(self.raw_l.astype(np.float32), axis=0, dtype=np.float32),
                    np.mean(self.raw_l, axis=0)),
                'adjusted_loss': adj_l,
                'adjusted_grad': np.mean(np.abs(adj_g), axis=1),
                'grad_diff': np.mean(np.abs(self.raw_g[0] - self.raw_g[1]), axis=1),
            }
        if advance and cfg.env == 'optlrs' and self.stream is not None \
                and self.ext_idx is None:
            self.stream.advance(everyone)                     # multioptlrs.py:128
        return obs, reward, done, info


class OptVecEnvOracle:
    """VecEnv surface over the batched oracle (vectorize/optvecenv.py:57-91): rows in
    lexicographic agent order, per-env reward/done replicated P times, auto-reset of
    finished envs with the reset observation returned (concurrentvecenv.py:32-38)."""

    def __init__(self, env: BatchedOptEnvOracle, lexicographic=True):
        self.env = env
        num_params = env.num_params
        self.perm = (lexicographic_rows(num_params) if lexicographic
                     else np.arange(num_params))
        self.num_envs = env.num_envs * num_params
        self.agent_no_list = [num_params] * env.num_envs

    def _rows(self, obs):
        return obs[:, self.perm, :].reshape(self.num_envs, -1)

    def reset(self, init_params=None):
        return self._rows(self.env.reset(init_params=init_params))

    def step(self, actions, reset_params=None):
        env = self.env
        rows = np.asarray(actions, np.float32).reshape(env.num_envs, env.num_params)
        natural = np.empty_like(rows)
        natural[:, self.perm] = rows
        obs, reward, done, info = env.step(natural)
        steps = env.current_step.copy()
        if done.any():
            obs = env.reset(done, init_params=reset_params)
        num_params = env.num_params
        info = dict(info, episode_r=reward, episode_l=steps)
        return (self._rows(obs), np.repeat(reward, num_params),
                np.repeat(done, num_params), info)

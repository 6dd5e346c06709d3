"""One-step checker (TEST INFRASTRUCTURE, like everything under oracle/): given the device's state of a
few sampled envs before a step and its outputs after it, replay the step with the oracle
(optenv_oracle.BatchedOptEnvOracle, the numpy restatement of envs/multioptlrs.py:80-129) and
compare.  Used by tests/ and by bench.py's ``parity`` entry, never by the product path.

The oracle is seeded with exactly the state the step reads (multioptlrs.py:85-101): w_{t-1},
g_{t-1}, L_{t-1}, the step counter and the minibatch.  That determines the new weights, loss,
reward, done and the NEWEST column of each observation block; the older columns are, by the ring
semantics of ``History`` (utils_common.py:129-137), the previous observation shifted by one, which is
checked bit for bit against the device's own previous observation.
"""
import numpy as np

from oracle import optenv_oracle as orc

RTOL = 1e-5


def replay_step(spec, feats, labels, state, actions_nat, max_batches=400, depth=5):
    """state: dict of arrays for n envs -- params [n,P] f32, grad_prev [n,P] f32, loss_prev [n],
    step [n] int, idx [n,B] int32, cnt [n] int32.  -> (oracle, obs, reward, done, info)."""
    n = state['params'].shape[0]
    ref = orc.BatchedOptEnvOracle(spec, feats, labels, n, batch_size=state['idx'].shape[1],
                                  config=orc.EnvConfig.multioptlrs(max_batches, depth),
                                  perms=np.zeros((n, len(feats)), np.int64))
    ref.weights = state['params'].astype(np.float32).copy()
    ref.raw_w[0] = ref.weights
    ref.raw_g[0] = state['grad_prev'].astype(np.float64)
    ref.raw_l[0] = np.asarray(state['loss_prev'], np.float64)
    ref.current_step[:] = state['step']
    ref.raw_pushes[:] = np.asarray(state['step']) + 1
    ref.set_batch(state['idx'], state['cnt'])
    obs, reward, done, info = ref.step(actions_nat)
    return ref, obs, reward, done, info


def compare_step(ref, ref_obs, ref_reward, ref_done, state, device, depth=5):
    """device: dict -- params [n,P], loss [n], reward [n], done [n] bool, obs [n,P,3H] and obs_prev
    [n,P,3H] in NATURAL parameter order.  Returns a dict of error statistics; 'ok' is the verdict
    at the bar of tests/test_gpu_parity.py (1e-5 relative on 98 % of the well-conditioned entries,
    20e-5 on all of them, ill-conditioned ratios after the clip within 1e-2)."""
    new_w, new_g = ref.weights.astype(np.float64), ref.raw_g[0]
    prev_w, prev_g = state['params'].astype(np.float64), state['grad_prev'].astype(np.float64)
    newest = [0, depth, 2 * depth]
    err = np.abs(device['obs'][:, :, newest] - ref_obs[:, :, newest]) / np.maximum(1.0, np.abs(ref_obs[:, :, newest] + 1.0))
    g_rms = np.sqrt(np.mean(new_g ** 2, axis=1, keepdims=True)) + 1e-30
    w_rms = np.sqrt(np.mean(prev_w ** 2, axis=1, keepdims=True)) + 1e-30
    ill = np.zeros(err.shape, bool)
    ill[:, :, 0] = np.maximum(np.abs(prev_w), np.abs(new_w)) < 1e-2 * w_rms
    ill[:, :, 2] = np.maximum(np.abs(prev_g), np.abs(new_g)) < 1e-2 * g_rms
    good = err[~ill]
    werr = np.abs(device['params'] - ref.weights) / np.maximum(np.abs(ref.weights), np.abs(ref.weights).mean())
    lerr = np.abs(device['loss'] - ref.raw_l[0]) / np.maximum(np.abs(ref.raw_l[0]), 1e-30)
    rerr = np.abs(device['reward'] - ref_reward) / np.maximum(np.abs(ref_reward), 1.0)
    keep = [c for c in range(3 * depth) if c % depth != depth - 1]
    shifted = [c + 1 for c in keep]
    shift_exact = bool(np.array_equal(device['obs'][:, :, shifted], device['obs_prev'][:, :, keep]))
    stats = {
        'tol': RTOL,
        'max_rel_err': float(max(good.max(), werr.max(), lerr.max(), rerr.max())),
        'obs_frac_within_tol': float(np.mean(good <= RTOL)),
        'obs_max_rel_err': float(good.max()),
        'ill_conditioned_max_abs_err': float(err[ill].max()) if ill.any() else 0.0,
        'weights_max_rel_err': float(werr.max()),
        'loss_max_rel_err': float(lerr.max()),
        'reward_max_rel_err': float(rerr.max()),
        'done_equal': bool(np.array_equal(np.asarray(device['done'], bool), ref_done)),
        'history_shift_exact': shift_exact,
    }
    stats['ok'] = bool(stats['done_equal'] and shift_exact and good.max() <= 20 * RTOL and
                       stats['obs_frac_within_tol'] > 0.98 and stats['ill_conditioned_max_abs_err'] <= 1e-2 and
                       werr.max() <= RTOL and lerr.max() <= RTOL and rerr.max() <= RTOL)
    return stats

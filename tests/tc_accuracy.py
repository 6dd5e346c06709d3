"""Ad-hoc probe (not a test): loss / gradient error of the eval kernels against the float64 oracle on
the config-4 shape, for B2E_TC=0 (FFMA) and B2E_TC=2 (tcgen05 3xTF32), as a fraction of the natural
scale (mean |g| per env) and of the sum of absolute terms of each dot product."""
import os
import sys

import numpy as np

sys.path.insert(0, '.')
from oracle import optenv_oracle as orc  # noqa: E402

spec = orc.ProblemSpec('softmax', 784, (64,), 10)
num_rows, batch, num_envs = 2048, 32, int(sys.argv[1]) if len(sys.argv) > 1 else 64
rng = np.random.RandomState(0)
feats = rng.uniform(size=(num_rows, 784)).astype(np.float32)
labels = rng.randint(0, 10, num_rows).astype(np.int32)
params = np.stack([orc.glorot_uniform_init(spec, rng) for _ in range(num_envs)])
params += 0.05 * rng.normal(size=params.shape).astype(np.float32)
idx = rng.randint(0, num_rows, size=(num_envs, batch)).astype(np.int32)
cnt = np.full(num_envs, batch, np.int32)
cnt[-1] = 11
mask = np.arange(batch)[None, :] < cnt[:, None]
ref_g, ref_l, terms = orc.loss_and_grad(spec, params, feats[idx], labels[idx], mask, abs_terms=True)
scale = np.abs(ref_g).mean(axis=1, keepdims=True)
for mode in ('0', '2'):
    os.environ['B2E_TC'] = mode
    os.environ['B2E_TC_CHECK'] = '1'
    from custom_envs_b200.batched_env import BatchedOptEnv, ProblemSpec
    env = BatchedOptEnv(ProblemSpec('softmax', 784, (64,), 10), feats, labels, num_envs, batch_size=batch,
                        index_mode='external', auto_reset=False)
    env.set_state('params', params)
    grad, loss = env.evaluate(idx, cnt)
    grad, loss = grad.cpu().numpy().astype(np.float64), loss.cpu().numpy().astype(np.float64)
    err = np.abs(grad - ref_g) / np.maximum(np.abs(ref_g), scale)
    w1 = slice(0, 784 * 64)
    noise = np.abs(grad - ref_g) / np.maximum(terms, 1e-30)
    print('B2E_TC=%s  loss rel err max %.2e | grad err / max(|g|, mean|g|): max %.2e  p99.9 %.2e  mean %.2e | W1 part max %.2e, tail part max %.2e'
          % (mode, np.max(np.abs(loss - ref_l) / np.abs(ref_l)), err.max(), np.quantile(err, 0.999), err.mean(),
             err[:, w1].max(), err[:, 784 * 64:].max()), flush=True)
    print('          |error| / sum|terms| of the element: max %.2e  p99.99 %.2e  p99.9 %.2e  mean %.2e' % (
        noise.max(), np.quantile(noise, 0.9999), np.quantile(noise, 0.999), noise.mean()), flush=True)
    env.close()

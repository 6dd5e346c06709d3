"""Ad-hoc probe (not a test): per-kernel ms of the large-problem step for each observation
kernel variant (B2E_OBS), and bit-equality of their outputs with the baseline kernel."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from custom_envs_b200.batched_env import BatchedOptEnv, ProblemSpec  # noqa: E402

parser = argparse.ArgumentParser()
parser.add_argument('--envs', type=int, default=2048)
parser.add_argument('--steps', type=int, default=12)
parser.add_argument('--variants', default='0,r1s,r1b,r2s,r2b,r4s,r4b,3s,3b,4s,4b,6s,6b')
parser.add_argument('--row-order', default='lexicographic')
parser.add_argument('--hidden', default='64', help='comma separated hidden widths, empty for none')
args = parser.parse_args()

rng = np.random.RandomState(0)
hidden = tuple(int(h) for h in args.hidden.split(',') if h)
spec, rows = ProblemSpec('softmax', 784, hidden, 10), 60000
feats = rng.uniform(size=(rows, spec.num_features)).astype(np.float32)
labels = rng.randint(0, spec.num_outputs, rows).astype(np.int32)
perm = np.arange(rows, dtype=np.int32)
rng.shuffle(perm)
ref = None
for variant in args.variants.split(','):
    os.environ['B2E_OBS'] = variant
    env = BatchedOptEnv(spec, feats, labels, args.envs, perms=perm, row_order=args.row_order, init_seed=3)
    env.reset()
    gen = torch.Generator(device=env.device)
    gen.manual_seed(5)
    actions = torch.rand(env.num_rows, device=env.device, generator=gen) * 3
    env.set_trace(True)
    acc = {}
    for i in range(args.steps):
        obs, reward, done, info = env.step(actions)
        if i >= 7:          # steady state: every ring slot valid
            for k, v in env.last_step_kernel_ms().items():
                acc.setdefault(k, []).append(v)
    torch.cuda.synchronize()
    check = (obs[::997].clone(), info.clone(), reward.clone())
    if ref is None:
        ref = check
        same = 'reference'
    else:
        same = 'obs_equal=%s info_close=%s reward_equal=%s' % (
            torch.equal(ref[0], check[0]),
            torch.allclose(ref[1], check[1], rtol=1e-9, atol=0, equal_nan=True),
            torch.equal(ref[2], check[2]))
    ms = {k: float(np.mean(v)) for k, v in acc.items()}
    obs_bytes = 26 * 4 * spec.size * args.envs
    print(f'{variant:5s} ' + ' '.join(f'{k}={v:.3f}' for k, v in ms.items())
          + f' | obs {obs_bytes / ms["obs_kernel"] / 1e6:7.1f} GB/s | total {sum(ms.values()):.3f} ms | {same}', flush=True)
    env.close()
    del env, obs, reward, done, info
    torch.cuda.empty_cache()

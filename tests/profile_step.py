"""Small driver for ncu captures (not a test): a few steps of the BASELINE config-4 shape."""
import argparse
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from custom_envs_b200.batched_env import BatchedOptEnv, ProblemSpec  # noqa: E402

parser = argparse.ArgumentParser()
parser.add_argument('--envs', type=int, default=592)
parser.add_argument('--steps', type=int, default=3)
parser.add_argument('--row-order', default='lexicographic')
parser.add_argument('--config', default='mlp', choices=['mlp', 'softmax', 'iris'])
args = parser.parse_args()
rng = np.random.RandomState(0)
if args.config == 'iris':
    spec, rows = ProblemSpec('softmax', 4, (), 3), 150
elif args.config == 'softmax':
    spec, rows = ProblemSpec('softmax', 784, (), 10), 60000
else:
    spec, rows = ProblemSpec('softmax', 784, (64,), 10), 60000
feats = rng.uniform(size=(rows, spec.num_features)).astype(np.float32)
labels = rng.randint(0, spec.num_outputs, rows).astype(np.int32)
perm = np.arange(rows, dtype=np.int32)
rng.shuffle(perm)
env = BatchedOptEnv(spec, feats, labels, args.envs, perms=perm, row_order=args.row_order)
env.reset()
actions = torch.rand(env.num_rows, device=env.device) * 3
start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(args.steps + 2):
    if i == 2:
        start.record()
    env.step(actions)
stop.record()
torch.cuda.synchronize()
print('ms/step', start.elapsed_time(stop) / args.steps, 'envs', args.envs)
env.close()

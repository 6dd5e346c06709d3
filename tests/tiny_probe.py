"""Ad-hoc probe (not a test): a few batched steps of BASELINE config 2 (iris-shaped softmax regression, 1024 envs), for ncu."""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from custom_envs_b200.batched_env import BatchedOptEnv, ProblemSpec  # noqa: E402

rng = np.random.RandomState(0)
feats = rng.uniform(size=(150, 4)).astype(np.float32)
labels = rng.randint(0, 3, 150).astype(np.int32)
perm = np.arange(150, dtype=np.int32)
rng.shuffle(perm)
env = BatchedOptEnv(ProblemSpec('softmax', 4, (), 3), feats, labels, 1024, perms=perm)
env.reset()
actions = torch.rand(env.num_rows, device=env.device) * 3
for _ in range(8):
    env.step(actions)
torch.cuda.synchronize()
print('ok')

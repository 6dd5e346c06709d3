"""Ad-hoc probe (not a test): a few batched steps of BASELINE config 3 (softmax regression 784 -> 10, 1024 envs), for ncu."""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from custom_envs_b200.batched_env import BatchedOptEnv, ProblemSpec  # noqa: E402

rng = np.random.RandomState(0)
rows = 60000
feats = rng.uniform(size=(rows, 784)).astype(np.float32)
labels = rng.randint(0, 10, rows).astype(np.int32)
perm = np.arange(rows, dtype=np.int32)
rng.shuffle(perm)
envs = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
env = BatchedOptEnv(ProblemSpec('softmax', 784, (), 10), feats, labels, envs, perms=perm)
env.reset()
actions = torch.rand(env.num_rows, device=env.device) * 3
for _ in range(6):
    env.step(actions)
env.set_trace(True)
env.step(actions)
print(env.last_step_kernel_ms())
torch.cuda.synchronize()

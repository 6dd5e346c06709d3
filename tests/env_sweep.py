"""Ad-hoc probe (not a test): BASELINE config 5 on one GPU -- env-steps/s of the MLP config
(784-64-10, B=32, H=5, lexicographic rows) for growing env counts."""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from custom_envs_b200.batched_env import BatchedOptEnv, ProblemSpec  # noqa: E402

spec, rows = ProblemSpec('softmax', 784, (64,), 10), 60000
rng = np.random.RandomState(0)
feats = rng.uniform(size=(rows, 784)).astype(np.float32)
labels = rng.randint(0, 10, rows).astype(np.int32)
perm = np.arange(rows, dtype=np.int32)
rng.shuffle(perm)
bytes_per_env_step = 4 * (spec.size * 28 + 32 * 785)
for envs in [int(a) for a in sys.argv[1:]] or [1024, 2048, 4096, 8192, 16384]:
    env = BatchedOptEnv(spec, feats, labels, envs, perms=perm)
    env.reset()
    actions = torch.rand(env.num_rows, device=env.device) * 3
    for _ in range(6):
        env.step(actions)
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 10
    start.record()
    for _ in range(steps):
        env.step(actions)
    stop.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(stop) / steps
    print(f'envs {envs:6d}  {ms:8.3f} ms/step  {envs / ms * 1e3:10.0f} env-steps/s  '
          f'{bytes_per_env_step * envs / ms / 1e6:7.1f} GB/s algorithmic '
          f'({bytes_per_env_step * envs / ms / 1e6 / 6550.7:.3f} of the measured HBM peak)', flush=True)
    env.close()
    del env, actions
    torch.cuda.empty_cache()

"""Ad-hoc probe (not a test): one full reset of the config-4 env batch, for an ncu launch list.
    python tests/reset_probe.py [envs]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                     # noqa: E402
from custom_envs_b200.batched_env import BatchedOptEnv, ProblemSpec, env_permutations_device   # noqa: E402

envs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
feats, labels = bench.synthetic_data()
env = BatchedOptEnv(ProblemSpec('softmax', bench.D, (bench.HID,), bench.C), feats, labels, envs, batch_size=32,
                    max_batches=400, max_history=5, perms=env_permutations_device(bench.ROWS, list(range(envs)), 'cuda:0'), init_seed=1)
env.reset()
start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
start.record()
env.reset()
stop.record()
torch.cuda.synchronize()
print('full reset of %d envs: %.3f ms' % (envs, start.elapsed_time(stop)))

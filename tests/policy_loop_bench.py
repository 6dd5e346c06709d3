"""Config 4 with a shared MLP policy in the loop, observations never leaving the GPU."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                     # noqa: E402
from custom_envs_b200.batched_env import BatchedOptEnv, ProblemSpec, env_permutations   # noqa: E402
from custom_envs_b200.vectorize.device_rollout import SharedMlpPolicy, device_rollout   # noqa: E402

envs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
feats, labels = bench.synthetic_data()
env = BatchedOptEnv(ProblemSpec('softmax', bench.D, (bench.HID,), bench.C), feats, labels, envs,
                    batch_size=32, max_batches=400, max_history=5,
                    perms=env_permutations(bench.ROWS, list(range(envs))), init_seed=1)
env.reset()
torch.manual_seed(0)
for name, dtype, tf32 in (('fp32', torch.float32, False), ('tf32', torch.float32, True), ('bf16', torch.bfloat16, False)):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    policy = SharedMlpPolicy(env.obs_dim).to(env.device, dtype)
    act = (lambda o: policy.act(o.to(dtype))) if dtype != torch.float32 else policy.act
    device_rollout(env, act, 2)
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    start.record()
    _, finished = device_rollout(env, act, steps)
    stop.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(stop) / steps
    print('policy %s: %.2f ms/step, %.0f env-steps/s (%d envs, %d agent rows, %d episodes ended)'
          % (name, ms, envs / ms * 1e3, envs, env.num_rows, finished), flush=True)

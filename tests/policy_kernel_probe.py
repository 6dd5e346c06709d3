"""Ad-hoc probe (not a test): the policy kernel alone on a dense observation matrix, for ncu.
    python tests/policy_kernel_probe.py [rows] [tanh_mode] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from custom_envs_b200.vectorize.device_policy import DevicePolicy   # noqa: E402
from custom_envs_b200.vectorize.device_rollout import SharedMlpPolicy   # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 512 * 50890
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
torch.manual_seed(0)
policy = SharedMlpPolicy(15).cuda()
obs = torch.randn(rows, 15, device='cuda')
out = torch.empty(rows, device='cuda')
dev = DevicePolicy.from_torch(policy.pi, tanh_mode=mode)
dev.act(obs, out)
start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
start.record()
for _ in range(reps):
    dev.act(obs, out)
stop.record()
torch.cuda.synchronize()
ms = start.elapsed_time(stop) / reps
print('policy kernel, %d rows, tanh mode %d: %.3f ms, %.0f G tanh/s, %.1f GB/s of observation rows'
      % (rows, mode, ms, rows * 128 / ms / 1e6, rows * 64 / ms / 1e6))

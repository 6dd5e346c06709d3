"""Ad-hoc probe (not a test): BASELINE config 5 -- the MLP config (784-64-10, B=32, H=5, lexicographic
rows, 60000 synthetic rows) for E_total envs sharded by index over the ranks of one box.

    python tests/scaling_sweep.py 1024 2048 ...                               # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29544 tests/scaling_sweep.py 1024 ... 65536             # eight ranks

Every env starts from its own rotation of one data-set permutation (distinct minibatches per env without
E host-side shuffles); the epoch permutation is shared.  Device-timed (CUDA events, barrier on both sides,
max over ranks); no collective on the step path.  Points whose per-GPU state would not fit are N/A."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, '.')
from custom_envs_b200.batched_env import BatchedOptEnv, ProblemSpec  # noqa: E402
from custom_envs_b200.sharding import shard_range  # noqa: E402

world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
device = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=device)
spec, rows = ProblemSpec('softmax', 784, (64,), 10), 60000
rng = np.random.RandomState(0)
feats = rng.uniform(size=(rows, 784)).astype(np.float32)
labels = rng.randint(0, 10, rows).astype(np.int32)
perm = np.arange(rows, dtype=np.int32)
rng.shuffle(perm)
bytes_per_env_step = 4 * (spec.size * 28 + 32 * 785)
bytes_of_state_per_env = 4 * (spec.size * (3 + 2 * 5 + 15) + 3 * rows)      # w, g x2, rings, observations, order x2 + init
MAX_ENVS_PER_GPU = 20480
steps, warmup = 8, 5
for total in [int(a) for a in sys.argv[1:]] or [1024, 2048, 4096, 8192, 16384]:
    first, envs = shard_range(total, world, rank)
    if -(-total // world) > MAX_ENVS_PER_GPU:
        if rank == 0:
            print(f'envs_total {total:6d} gpus {world}  N/A: {-(-total // world)} envs per GPU need '
                  f'{bytes_of_state_per_env * -(-total // world) / 1e9:.0f} GB of HBM', flush=True)
        continue
    base = torch.as_tensor(perm.astype(np.int64), device=device)
    ar = torch.arange(rows, device=device)[None, :]
    shift = (torch.arange(first, first + envs, device=device)[:, None] * 977) % rows
    init_orders = base[(ar + shift) % rows].to(torch.int32)
    env = BatchedOptEnv(spec, feats, labels, envs, perms=perm, init_orders=init_orders, device=device, init_seed=7 + rank)
    del init_orders, shift
    env.reset()
    actions = torch.rand(env.num_rows, device=device) * 3
    for _ in range(warmup):
        env.step(actions)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(steps):
        env.step(actions)
    stop.record()
    torch.cuda.synchronize()
    ms = torch.tensor([start.elapsed_time(stop) / steps], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    env.set_trace(True)
    env.step(actions)
    kernel_ms = env.last_step_kernel_ms()
    if rank == 0:
        ms = float(ms.item())
        print(f'envs_total {total:6d} gpus {world} envs_per_gpu {envs:6d}  {ms:8.3f} ms/step  {total / ms * 1e3:11.0f} env-steps/s  '
              f'{bytes_per_env_step * total / ms / 1e6 / world:7.1f} GB/s per GPU algorithmic '
              f'({bytes_per_env_step * total / ms / 1e6 / world / 6550.7:.3f} of the measured HBM peak)  kernels(ms) '
              + ' '.join('%s=%.3f' % (k.replace('eval_kernel', 'eval'), v) for k, v in kernel_ms.items()), flush=True)
    env.close()
    del env, actions
    torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()

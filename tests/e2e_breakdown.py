"""Phase timing of DeviceOptVecEnv.step at config 4 (host buffers): staging, H2D, step, D2H."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                     # noqa: E402
from custom_envs_b200.batched_env import BatchedOptEnv, ProblemSpec, env_permutations   # noqa: E402
from custom_envs_b200.vectorize import optvecenv as ov          # noqa: E402

envs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
feats, labels = bench.synthetic_data()
env = BatchedOptEnv(ProblemSpec('softmax', bench.D, (bench.HID,), bench.C), feats, labels, envs,
                    batch_size=32, max_batches=400, max_history=5,
                    perms=env_permutations(bench.ROWS, list(range(envs))), init_seed=1)
env.reset()
vec = ov.DeviceOptVecEnv(env)
acts = np.random.RandomState(3).uniform(0, 3, size=(env.num_rows, 1)).astype(np.float32)
vec.step(acts)
sync = torch.cuda.synchronize
for rep in range(3):
    sync()
    t = [time.perf_counter()]
    flat = np.ascontiguousarray(np.asarray(acts, np.float32).reshape(-1)); t.append(time.perf_counter())
    ov.stage_to_device(flat, vec._act_host, vec._act_dev); t.append(time.perf_counter())
    sync(); t.append(time.perf_counter())
    obs, reward, done, info = env.step(vec._act_dev); sync(); t.append(time.perf_counter())
    vec._rew_host.copy_(reward, non_blocking=True); vec._done_host.copy_(done, non_blocking=True)
    vec._info_host.copy_(info, non_blocking=True); sync(); t.append(time.perf_counter())
    vec._expand_rows(); t.append(time.perf_counter())
    vec._obs_host.copy_(obs, non_blocking=True); sync(); t.append(time.perf_counter())
    names = ['asarray', 'stage+queue', 'h2d tail', 'step', 'scalars d2h', 'expand rows (alone)', 'obs d2h']
    print(' | '.join('%s %.1f ms' % (n, 1e3 * (b - a)) for n, a, b in zip(names, t, t[1:])), flush=True)
    t0 = time.perf_counter(); vec.step(acts); sync(); print('whole step %.1f ms' % (1e3 * (time.perf_counter() - t0)))
    t0 = time.perf_counter(); vec.step_async(acts); t1 = time.perf_counter(); vec.step_wait(); t2 = time.perf_counter()
    print('step_async %.1f ms, step_wait %.1f ms' % (1e3 * (t1 - t0), 1e3 * (t2 - t1)), flush=True)
# obs copy in chunks (does the copy engine care?)
for chunks in (1, 4, 16):
    rows = env.num_rows
    cuts = np.linspace(0, rows, chunks + 1).astype(np.int64)
    sync(); t0 = time.perf_counter()
    for lo, hi in zip(cuts, cuts[1:]):
        vec._obs_host[lo:hi].copy_(obs[lo:hi], non_blocking=True)
    sync(); dt = time.perf_counter() - t0
    print('obs d2h in %d chunks: %.1f ms, %.1f GB/s' % (chunks, 1e3 * dt, obs.numel() * 4 / dt / 1e9))

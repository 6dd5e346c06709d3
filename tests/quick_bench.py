"""Ad-hoc timing probe (not a test): ms/step of the fused kernel at the BASELINE shapes."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from custom_envs_b200.batched_env import BatchedOptEnv, ProblemSpec

def run(name, spec, num_rows, num_envs, order, steps=20):
    rng = np.random.RandomState(0)
    feats = rng.uniform(size=(num_rows, spec.num_features)).astype(np.float32)
    labels = rng.randint(0, spec.num_outputs, num_rows).astype(np.int32)
    perm = np.arange(num_rows, dtype=np.int32); rng.shuffle(perm)
    env = BatchedOptEnv(spec, feats, labels, num_envs, perms=perm, row_order=order)
    env.reset()
    actions = torch.rand(env.num_rows, device=env.device) * 3
    for _ in range(3):
        env.step(actions)
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        env.step(actions)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / steps
    P = spec.size
    bytes_step = 4 * (P * 28 + 32 * (spec.num_features + 1)) * num_envs
    print(f'{name:28s} E={num_envs:5d} {order:13s} {ms:9.3f} ms/step  {num_envs/ms*1e3:12.0f} env-steps/s  '
          f'{bytes_step/ms/1e6:8.1f} GB/s algorithmic', flush=True)
    env.close()

if __name__ == '__main__':
    import os
    os.environ['B2E_FORCE_GENERIC'] = '1'
    run('cfg3 softmax 784->10 GENERIC', ProblemSpec('softmax', 784, (), 10), 60000, 1024, 'lexicographic')
    run('mlp 784->256->256->10', ProblemSpec('softmax', 784, (256, 256), 10), 60000, 512, 'lexicographic', steps=5)
    del os.environ['B2E_FORCE_GENERIC']
    for order in ('lexicographic',):
        run('cfg2 softmax 4->3', ProblemSpec('softmax', 4, (), 3), 150, 1024, order)
        run('cfg3 softmax 784->10', ProblemSpec('softmax', 784, (), 10), 60000, 1024, order)
        run('cfg4 mlp 784->64->10', ProblemSpec('softmax', 784, (64,), 10), 60000, 592, order, steps=10)
        run('cfg4 mlp 784->64->10', ProblemSpec('softmax', 784, (64,), 10), 60000, 4096, order, steps=5)

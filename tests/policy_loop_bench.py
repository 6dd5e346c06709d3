"""Config 4 with a shared MLP policy in the loop, observations never leaving the GPU:
torch policies (fp32 / tf32 / bf16, chunked) against the library's tcgen05 policy kernel on observation
rows and on the env's rings (ring-only env steps).

    python tests/policy_loop_bench.py [envs] [steps] [torch|notorch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                     # noqa: E402
from custom_envs_b200.batched_env import BatchedOptEnv, ProblemSpec, env_permutations_device   # noqa: E402
from custom_envs_b200.vectorize.device_policy import DevicePolicy, device_policy_rollout   # noqa: E402
from custom_envs_b200.vectorize.device_rollout import SharedMlpPolicy, device_rollout   # noqa: E402

envs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
with_torch = (sys.argv[3] if len(sys.argv) > 3 else 'torch') == 'torch'
feats, labels = bench.synthetic_data()
perms = env_permutations_device(bench.ROWS, list(range(envs)), 'cuda:0')


def make_env(materialize_obs=True):
    env = BatchedOptEnv(ProblemSpec('softmax', bench.D, (bench.HID,), bench.C), feats, labels, envs,
                        batch_size=32, max_batches=400, max_history=5, perms=perms, init_seed=1,
                        materialize_obs=materialize_obs)
    env.reset()
    return env


def timed(fn, reps):
    fn()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    start.record()
    for _ in range(reps):
        out = fn()
    stop.record()
    torch.cuda.synchronize()
    return start.elapsed_time(stop) / reps, out


torch.manual_seed(0)
env = make_env()
policy32 = SharedMlpPolicy(env.obs_dim).to(env.device)
if with_torch:
    for name, dtype, tf32 in (('fp32', torch.float32, False), ('tf32', torch.float32, True), ('bf16', torch.bfloat16, False)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        policy = SharedMlpPolicy(env.obs_dim).to(env.device, dtype)
        act = (lambda o: policy.act(o.to(dtype))) if dtype != torch.float32 else policy.act
        device_rollout(env, act, 2)
        ms, (_, finished) = timed(lambda: device_rollout(env, act, steps), 1)
        ms /= steps
        print('torch policy %s: %.2f ms/step, %.0f env-steps/s (%d envs, %d agent rows, %d episodes ended)'
              % (name, ms, envs / ms * 1e3, envs, env.num_rows, finished), flush=True)

actions = torch.empty(env.num_rows, device=env.device)
for mode, tag in ((0, 'tanh.approx.f32'), (1, 'layer 1 tanh.approx.bf16x2'), (2, 'both layers tanh.approx.bf16x2')):
    dev = DevicePolicy.from_torch(policy32, tanh_mode=mode)
    ms_k, _ = timed(lambda: dev.act(env.obs, actions), 5)
    ms_r, _ = timed(lambda: dev.act_env(env, actions), 5)
    flops = 2.0 * env.num_rows * (16 * 64 + 80 * 64 + 64)
    print('policy kernel (%s): dense rows %.3f ms (%.0f TFLOP/s bf16, %.0f G tanh/s), rings %.3f ms'
          % (tag, ms_k, flops / ms_k / 1e9, env.num_rows * 128 / ms_k / 1e6, ms_r), flush=True)
    ms, (_, finished) = timed(lambda: device_policy_rollout(env, dev, steps, ring_only=False), 1)
    ms /= steps
    print('  rollout on observation rows: %.2f ms/step, %.0f env-steps/s' % (ms, envs / ms * 1e3), flush=True)
    ms, (_, finished) = timed(lambda: device_policy_rollout(env, dev, steps, ring_only=True), 1)
    ms /= steps
    print('  rollout on the rings (ring-only env steps): %.2f ms/step, %.0f env-steps/s' % (ms, envs / ms * 1e3), flush=True)
    if mode == 0:
        ms, (_, finished) = timed(lambda: device_policy_rollout(env, dev, steps + 2, ring_only=True, use_graph=True), 1)
        ms /= steps + 2
        print('  the same as a CUDA graph of two steps (capture included in the timing): %.2f ms/step, %.0f env-steps/s'
              % (ms, envs / ms * 1e3), flush=True)
    dev.close()
# the ring-only env step alone
gen_actions = torch.rand(env.num_rows, device=env.device) * 3
ms_d, _ = timed(lambda: env.step(gen_actions), 10)
ms_r, _ = timed(lambda: env.step(gen_actions, ring_only=True), 10)
env.set_trace(True)
env.step(gen_actions, ring_only=True)
print('env step alone: with observation rows %.3f ms, ring-only %.3f ms  kernels(ms) %s'
      % (ms_d, ms_r, ' '.join('%s=%.3f' % kv for kv in env.last_step_kernel_ms().items())), flush=True)

"""Ad-hoc probe (not a test): the launch-bound config 2 (softmax regression on iris-shaped data, 1024
envs, one fused kernel per step) eagerly, replayed from a CUDA graph, and next to the launch floor of
the same box (a graph of empty-handed kernels), plus config 3 / config 4 replayed two steps per graph."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from custom_envs_b200.batched_env import BatchedOptEnv, ProblemSpec  # noqa: E402


def timed(fn, reps):
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    start.record()
    for _ in range(reps):
        fn()
    stop.record()
    torch.cuda.synchronize()
    return start.elapsed_time(stop) * 1e3 / reps, (time.perf_counter() - t0) * 1e6 / reps     # device us, wall us


def floor(kernels, reps=200):
    """A graph of `kernels` dependent one-thread kernels: what a launch costs at best on this box."""
    x = torch.zeros(1, device='cuda')
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        x.add_(1.0)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(kernels):
            x.add_(1.0)
    graph.replay()
    dev, wall = timed(graph.replay, reps)
    return dev / kernels


def run(name, spec, rows, envs, steps_per_graph, reps):
    rng = np.random.RandomState(0)
    feats = rng.uniform(size=(rows, spec.num_features)).astype(np.float32)
    labels = rng.randint(0, spec.num_outputs, rows).astype(np.int32)
    perm = np.arange(rows, dtype=np.int32)
    rng.shuffle(perm)
    env = BatchedOptEnv(spec, feats, labels, envs, perms=perm)
    env.reset()
    actions = torch.rand(env.num_rows, device=env.device) * 3
    for _ in range(5):
        env.step(actions)
    eager_dev, eager_wall = timed(lambda: env.step(actions), reps)
    launches = env.launch_count
    env.step(actions)
    per_step = env.launch_count - launches
    out = '%-26s E=%5d  eager %8.1f us/step (host loop %8.1f us) %d launches/step' % (name, envs, eager_dev, eager_wall, per_step)
    for steps in steps_per_graph:
        graph, steps = env.capture_step_graph(actions, steps)
        graph.replay()
        dev, wall = timed(graph.replay, max(4, reps // steps))
        out += ' | graph of %2d steps: %8.1f us/step' % (steps, dev / steps)
        del graph
    print(out, flush=True)
    env.close()


if __name__ == '__main__':
    print('launch floor: %.2f us per kernel in a graph of 20 dependent one-thread kernels, %.2f us in a graph of 1'
          % (floor(20), floor(1)), flush=True)
    run('cfg2 softmax 4->3', ProblemSpec('softmax', 4, (), 3), 150, 1024, (1, 20), 400)
    run('cfg3 softmax 784->10', ProblemSpec('softmax', 784, (), 10), 60000, 1024, (2, 20), 100)
    run('cfg4 mlp 784->64->10', ProblemSpec('softmax', 784, (64,), 10), 60000, 1024, (2,), 20)

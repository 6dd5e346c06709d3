"""CPU baseline, path (i) of SURVEY 8d, with the REFERENCE'S OWN code: per-env `MultiOptLRs` objects (the reference's
`History`, `utils_env`, `BaseEnvironment`, per-agent dict building, 14 info statistics) driven by the reference's
`OptVecEnv` (one thread and one pipe per env, `vectorize/concurrentvecenv.py:27-104`, `optvecenv.py:17-91`), on the
BASELINE config-4 shape (MLP 784-64-10, P = 50 890 agents per env, minibatch 32).  Only the TensorFlow problem is
replaced (absent from the image): `NumpyProblem` of gen_golden.py, the oracle's float32 loss / gradient behind the
reference's `BaseProblem` interface -- it costs ~1 % of a step, the rest is the reference's Python.

Run in the build container only (needs /root/reference; nothing in tests/ or bench.py reads it at run time):

    python tests/golden/time_reference_faithful.py > profiles/r2_cpu_faithful_reference_code.txt
"""
import os
import sys
import time

import numpy as np
import numpy.random as npr

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as gg                                       # noqa: E402  (installs the stubs, imports the reference)

orc = gg.orc
spec = orc.ProblemSpec('softmax', 784, (64,), 10)
rows, batch = 4096, 32
data_rng = npr.RandomState(0)
feats = data_rng.uniform(size=(rows, 784)).astype(np.float32)
labels = np.arange(rows) % 10
cores = os.cpu_count() or 1
print('host: %d cores; reference tree: %s' % (cores, gg.custom_envs.__file__))
print('shape: MultiOptLRs over MLP 784-64-10 (P = %d agent rows per env), minibatch %d, %d-row data set' % (spec.size, batch, rows))
for num_envs, steps in ((1, 5), (min(4, cores), 3), (cores, 3)):
    envs = []
    for i in range(num_envs):
        npr.seed(1000 + i)
        problem = gg.NumpyProblem(spec, feats, labels, batch, npr.RandomState(100 + i), dict())
        gg.ref_multioptlrs.get_problem = lambda *a, problem=problem, **k: problem
        env = gg.ref_multioptlrs.MultiOptLRs(problem='nn', max_batches=400, max_history=5)
        env.seed(7 + i)
        envs.append(env)
    vec = gg.OptVecEnv([(lambda e=e: e) for e in envs])       # the reference's thread-per-env pipe vectoriser
    vec.reset()
    act_rng = npr.RandomState(2)
    actions = act_rng.uniform(0, 3, size=(num_envs * spec.size, 1)).astype(np.float32)
    vec.step(actions)                                          # warm-up
    t0 = time.perf_counter()
    for _ in range(steps):
        states, rewards, dones, infos = vec.step(actions)
    dt = time.perf_counter() - t0
    assert np.asarray(states).shape == (num_envs * spec.size, 15)
    print('%3d envs (one thread + pipe each), %d steps: %.2f s per batched step, %.3f env-steps/s'
          % (num_envs, steps, dt / steps, num_envs * steps / dt), flush=True)
    vec.close()
print('extrapolation to 4096 envs is linear in the env count (the threads share the interpreter lock).')

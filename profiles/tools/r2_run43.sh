#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_data_frontend.py -q -m gpu -k "device_shuffle" 2>&1 | tail -8
python - <<'PY'
import time, numpy as np, torch, sys
sys.path.insert(0, '.')
from custom_envs_b200.batched_env import env_permutations, env_permutations_device
for envs in (1024, 4096, 16384):
    env_permutations_device(60000, list(range(32)), 'cuda:0')
    torch.cuda.synchronize(); t0 = time.perf_counter()
    p = env_permutations_device(60000, list(range(envs)), 'cuda:0')
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    t0 = time.perf_counter(); env_permutations(60000, list(range(64))); host = (time.perf_counter() - t0) / 64 * envs
    print('permutations of 60000 rows for %5d envs: device %.3f s, numpy on the host %.2f s (extrapolated from 64)' % (envs, dt, host), flush=True)
PY

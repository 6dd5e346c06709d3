#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B2E_TC_CHECK=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_policy.py -q -m gpu -k "mlp_784x64x10 or cfg4 or reset_pipeline or full_size or ring_only or tc2" > gpurun_out/r2_tests_j.txt 2>&1
tail -5 gpurun_out/r2_tests_j.txt
timeout 300 python tests/obs_sweep.py --envs 4096 --variants r4b --steps 12 2>&1 | tail -2
timeout 300 python tests/tc_accuracy.py 64 2>&1 | tail -4

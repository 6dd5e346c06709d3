#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_policy.py -q -x -m gpu 2>&1 | tail -3
timeout 900 python tests/policy_loop_bench.py 4096 10 notorch > gpurun_out/r2_policy_loop_f.txt 2>&1
cat gpurun_out/r2_policy_loop_f.txt | tail -20

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_policy.py -q -x -m gpu > gpurun_out/r2_policy_tests_a.txt 2>&1
tail -25 gpurun_out/r2_policy_tests_a.txt

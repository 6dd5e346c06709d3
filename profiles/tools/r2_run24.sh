#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python tests/policy_loop_bench.py 4096 10 notorch > gpurun_out/r2_policy_loop_a.txt 2>&1
cat gpurun_out/r2_policy_loop_a.txt | tail -20

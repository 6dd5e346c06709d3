#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_policy.py -q -x -m gpu 2>&1 | tail -3
timeout 900 python tests/policy_loop_bench.py 4096 10 notorch > gpurun_out/r2_policy_loop_b.txt 2>&1
cat gpurun_out/r2_policy_loop_b.txt | tail -20
timeout 300 python tests/policy_kernel_probe.py 26055680 0 3 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:policy_kernel -s 1 -c 1 -o gpurun_out/r2_policy_v1 \
    python tests/policy_kernel_probe.py 26055680 0 1 > gpurun_out/r2_policy_v1_ncu.log 2>&1
tail -3 gpurun_out/r2_policy_v1_ncu.log

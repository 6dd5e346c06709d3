#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2_gpu_tests_c.txt 2>&1
tail -8 gpurun_out/r2_gpu_tests_c.txt
timeout 300 python tests/tc_accuracy.py 64 > gpurun_out/r2_tc_accuracy_c.txt 2>&1; cat gpurun_out/r2_tc_accuracy_c.txt
B2E_TC_TRACE=gpurun_out/r2_tc2_v15_trace.txt timeout 300 python tests/obs_sweep.py --envs 4096 --variants r4b --steps 12 > gpurun_out/r2_tc2_v15_timing.txt 2>&1; cat gpurun_out/r2_tc2_v15_timing.txt

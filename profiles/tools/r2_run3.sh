#!/bin/bash
# round 2, GPU call 3: tc2 v2 (dynamic MMA scheduling, L2 prefetch)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== tc2 parity" > gpurun_out/r2_run3_tests.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "tc2 or ffma" >> gpurun_out/r2_run3_tests.txt 2>&1
echo "== accuracy" > gpurun_out/r2_run3_accuracy.txt
timeout 300 python tests/tc_accuracy.py 64 >> gpurun_out/r2_run3_accuracy.txt 2>&1
echo "== timing" > gpurun_out/r2_run3_timing.txt
B2E_TC=2 timeout 300 python tests/obs_sweep.py --envs 4096 --variants r4b --steps 9 >> gpurun_out/r2_run3_timing.txt 2>&1 &&
B2E_TC=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc2_eval -s 2 -c 2 \
    -o gpurun_out/r2_tc2_v2 python tests/obs_sweep.py --envs 4096 --variants r4b --steps 9 > gpurun_out/r2_run3_ncu.log 2>&1
tail -n 5 gpurun_out/r2_run3_tests.txt gpurun_out/r2_run3_accuracy.txt gpurun_out/r2_run3_timing.txt gpurun_out/r2_run3_ncu.log

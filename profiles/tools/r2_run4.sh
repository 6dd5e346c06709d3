#!/bin/bash
# round 2, GPU call 4: tc2 v3 (batched converter loads, sleeping MMA poll)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T=${1:-v3}
echo "== tc2 parity" > gpurun_out/r2_tc2_${T}_tests.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "tc2 or ffma" >> gpurun_out/r2_tc2_${T}_tests.txt 2>&1
echo "== timing" > gpurun_out/r2_tc2_${T}_timing.txt
B2E_TC=2 timeout 300 python tests/obs_sweep.py --envs 4096 --variants r4b --steps 12 >> gpurun_out/r2_tc2_${T}_timing.txt 2>&1
B2E_TC=2 timeout 300 python tests/obs_sweep.py --envs 4096 --variants r4b --steps 9 >> gpurun_out/r2_tc2_${T}_timing.txt 2>&1 &&
B2E_TC=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc2_eval -s 2 -c 2 \
    -o gpurun_out/r2_tc2_${T} python tests/obs_sweep.py --envs 4096 --variants r4b --steps 9 > gpurun_out/r2_tc2_${T}_ncu.log 2>&1
tail -n 4 gpurun_out/r2_tc2_${T}_tests.txt gpurun_out/r2_tc2_${T}_timing.txt gpurun_out/r2_tc2_${T}_ncu.log

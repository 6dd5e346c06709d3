#!/bin/bash
# round 2, GPU call 1: descriptor conventions, tc2 parity, full-size sampled parity, eval timing
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_run1_gpu.txt
P=profiles/tools/bin/tc2_desc_probe
{
for kase in 0 1; do
  for v in "1 1 1 4" "1 1 0 4" "1 1 1 8" "1 1 0 8" "2 2 1 8" "2 2 0 8" "0 0 1 8" "0 0 0 8" "1 2 1 4" "2 1 1 8" "1 0 1 4" "2 2 1 4"; do
    timeout 60 $P $kase $v 2>&1 | tail -2
  done
done
} > gpurun_out/r2_desc_probe.txt 2>&1
echo "== tc2 parity" > gpurun_out/r2_run1_tests.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "tc2 or ffma" >> gpurun_out/r2_run1_tests.txt 2>&1
echo "== sampled parity at baseline sizes" >> gpurun_out/r2_run1_tests.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "sampled_env_parity" >> gpurun_out/r2_run1_tests.txt 2>&1
echo "== eval timing: default vs tc2" > gpurun_out/r2_run1_timing.txt
B2E_TC=0 timeout 300 python tests/obs_sweep.py --envs 4096 --variants r4b >> gpurun_out/r2_run1_timing.txt 2>&1
B2E_TC=2 B2E_TC_CHECK=1 timeout 300 python tests/obs_sweep.py --envs 4096 --variants r4b >> gpurun_out/r2_run1_timing.txt 2>&1
B2E_TC=2 timeout 300 python tests/obs_sweep.py --envs 4096 --variants r4b >> gpurun_out/r2_run1_timing.txt 2>&1
tail -5 gpurun_out/r2_desc_probe.txt gpurun_out/r2_run1_tests.txt gpurun_out/r2_run1_timing.txt

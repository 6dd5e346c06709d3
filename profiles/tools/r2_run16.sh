#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T=${1:-v11}
echo "== tc2 parity" > gpurun_out/r2_tc2_${T}_tests.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "tc2" >> gpurun_out/r2_tc2_${T}_tests.txt 2>&1
for PF in 1 0; do
  echo "== timing, B2E_TC_PF=$PF" >> gpurun_out/r2_tc2_${T}_timing.txt
  B2E_TC_PF=$PF B2E_TC=2 B2E_TC_TRACE=gpurun_out/r2_tc2_${T}_pf${PF}_trace.txt timeout 300 python tests/obs_sweep.py --envs 4096 --variants r4b --steps 12 >> gpurun_out/r2_tc2_${T}_timing.txt 2>&1
done
grep -E "passed|failed|==" gpurun_out/r2_tc2_${T}_tests.txt; cat gpurun_out/r2_tc2_${T}_timing.txt

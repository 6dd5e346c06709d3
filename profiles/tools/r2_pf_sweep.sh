#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out/r2_tc2_pf_sweep.txt
: > $O
for W in 1 0; do for PF in 0 3 10 25 50; do
  echo "== WTMA=$W PF=$PF" >> $O
  B2E_TC_WTMA=$W B2E_TC_PF=$PF B2E_TC=2 timeout 300 python tests/obs_sweep.py --envs 4096 --variants r4b --steps 10 2>&1 | tail -1 >> $O
done; done
cat $O

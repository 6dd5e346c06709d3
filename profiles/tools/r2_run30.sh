#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tests/reset_probe.py 4096 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"reset_|tc2_eval" --csv --log-file gpurun_out/r2_reset_launches.csv python tests/reset_probe.py 4096 > /dev/null 2>&1
grep -v "^==" gpurun_out/r2_reset_launches.csv | awk -F'","' '{print $5, $NF}' | tail -12

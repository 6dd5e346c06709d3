#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:thin3_eval -s 8 -c 1 -o gpurun_out/r2_thin3_v1 python tests/thin_probe.py > gpurun_out/r2_thin3_v1_ncu.log 2>&1
tail -2 gpurun_out/r2_thin3_v1_ncu.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tests/profile_step.py --envs 4096 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"obs_kernel3|update_kernel" -s 6 -c 2 -o gpurun_out/r2_obs_update python tests/profile_step.py --envs 4096 > gpurun_out/r2_obs_update_ncu.log 2>&1
tail -2 gpurun_out/r2_obs_update_ncu.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -q -m gpu -k "warp_per_env" > gpurun_out/r2_tests_h.txt 2>&1
tail -15 gpurun_out/r2_tests_h.txt

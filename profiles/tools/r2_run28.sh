#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "reset_pipeline" > gpurun_out/r2_tests_g.txt 2>&1
tail -30 gpurun_out/r2_tests_g.txt

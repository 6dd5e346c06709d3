#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dropin.py -q -m gpu -k "bench_line" 2>&1 | tail -12

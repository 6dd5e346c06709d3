#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_policy.py -q -m gpu -k "ring_only_step_is" 2>&1 | tail -12

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 600 python -m pytest tests/test_gpu_policy.py tests/test_data_frontend.py -q -m gpu -k "not_current or device_shuffle" 2>&1 | tail -5

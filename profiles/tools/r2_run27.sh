#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
rm -f gpurun_out/r2_test_stats.jsonl
B2E_TEST_STATS=gpurun_out/r2_test_stats.jsonl timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "golden or multioptimize" > gpurun_out/r2_tests_e.txt 2>&1
tail -3 gpurun_out/r2_tests_e.txt

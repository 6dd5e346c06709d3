#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2_gpu_tests_e.txt 2>&1
tail -8 gpurun_out/r2_gpu_tests_e.txt
timeout 900 python bench.py > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; tail -c 600 gpurun_out/r2_bench_d.json

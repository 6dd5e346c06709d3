#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tests/graph_bench.py > gpurun_out/r2_graph_bench_d.txt 2>&1; cat gpurun_out/r2_graph_bench_d.txt
timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_policy.py -q -m gpu 2>&1 | tail -3

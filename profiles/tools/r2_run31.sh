#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -q -m gpu -k "warp_per_env or iris or linreg or golden or index_stream or graph or script or optvecenv or env_step" > gpurun_out/r2_tests_h.txt 2>&1
tail -4 gpurun_out/r2_tests_h.txt
timeout 300 python tests/graph_bench.py > gpurun_out/r2_graph_bench_c.txt 2>&1; head -2 gpurun_out/r2_graph_bench_c.txt

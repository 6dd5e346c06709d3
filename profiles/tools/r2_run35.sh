#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_policy.py -q -m gpu -k "one_launch_step or resident_minibatch or softmax_784 or cfg3 or sampled" > gpurun_out/r2_tests_i.txt 2>&1
tail -12 gpurun_out/r2_tests_i.txt
timeout 300 python tests/thin_probe.py
timeout 300 python tests/graph_bench.py 2>&1 | grep cfg3

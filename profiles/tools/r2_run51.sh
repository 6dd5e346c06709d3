#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 300 python tests/thin_probe.py 2>&1 | tail -1
timeout 300 python tests/graph_bench.py 2>&1 | grep "cfg3\|cfg4"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "softmax_784 or cfg3 or sampled or mlp_784x64x10 or full_size" 2>&1 | tail -2
timeout 300 python tests/scaling_sweep.py 1024 4096 2>&1 | tail -2 | cut -c1-260

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_policy.py -q -m gpu -k "not_current" 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus 2 --skip-faithful > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_2gpu.json'))
print(d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['policy_loop'])
PY

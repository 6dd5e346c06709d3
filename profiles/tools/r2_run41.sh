#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-8}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus $N --steps 50 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
echo "rc $?"; wc -l gpurun_out/r2_bench_${N}gpu.json; tail -3 gpurun_out/r2_bench_${N}gpu.err | cut -c1-300
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_${N}gpu.json'))
print(d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['policy_loop'], d['parity']['ok'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29549 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | cut -c1-200

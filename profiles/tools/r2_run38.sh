#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1; tail -3 gpurun_out/r2_smoke.txt
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2_gpu_tests_f.txt 2>&1
tail -4 gpurun_out/r2_gpu_tests_f.txt
timeout 900 python bench.py > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_e.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['policy_loop'], d['parity']['ok'], d['cpu_baseline']['value'])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-400

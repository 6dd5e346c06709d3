#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2_gpu_tests_i.txt 2>&1
tail -3 gpurun_out/r2_gpu_tests_i.txt
timeout 900 python bench.py --skip-faithful > gpurun_out/r2_bench_h.json 2> gpurun_out/r2_bench_h.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_h.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['policy_loop']['value'], d['parity']['ok'])
PY

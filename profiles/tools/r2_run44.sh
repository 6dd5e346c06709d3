#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2_gpu_tests_h.txt 2>&1
tail -3 gpurun_out/r2_gpu_tests_h.txt
timeout 900 python bench.py --skip-faithful > gpurun_out/r2_bench_g.json 2> gpurun_out/r2_bench_g.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_g.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['policy_loop']['value'], d['parity']['ok'])
PY

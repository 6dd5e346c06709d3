#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python bench.py --skip-faithful > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_c.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['config']['step_ms_in_window'], d['config']['full_reset_ms'], d['parity']['ok'])
PY

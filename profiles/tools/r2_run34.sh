#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 3 --e2e-steps 1 --skip-faithful > gpurun_out/r2_bench_short.json 2> gpurun_out/r2_bench_short.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --e2e-steps 1 --skip-faithful > gpurun_out/r2_launches_bench.log 2>&1
grep -c "" gpurun_out/r2_launches_bench.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc2_eval -s 4 -c 2 -o gpurun_out/r2_tc2_v16 \
    python tests/profile_step.py --envs 4096 > gpurun_out/r2_tc2_v16_ncu.log 2>&1
tail -2 gpurun_out/r2_tc2_v16_ncu.log

#!/bin/bash
# config 5 sweep on N ranks + the N-rank host<->device copy probe (platform ceiling of the e2e path)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-8}
bash profiles/tools/r2_sweep.sh $N
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 \
    tests/pcie_probe.py > gpurun_out/r2_pcie_probe_${N}gpu.txt 2>&1
grep "d2h\|h2d" gpurun_out/r2_pcie_probe_${N}gpu.txt | sort | head -20

#!/bin/bash
# config 5: env-count sweep on N ranks of one box (N = number of GPUs this call was given)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-1}
O=gpurun_out/r2_cfg5_sweep_${N}gpu.txt
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv,noheader > ${O}.gpus
if [ "$N" = "1" ]; then
  timeout 1200 python tests/scaling_sweep.py 1024 2048 4096 8192 16384 32768 65536 > $O 2>&1
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
      tests/scaling_sweep.py 1024 2048 4096 8192 16384 32768 65536 > $O 2>&1
fi
grep envs_total $O

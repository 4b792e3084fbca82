#!/bin/bash
# the driver's multi-GPU command (default weak line + extra legs) with a tight timeout
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
(time timeout 280 $TR bench.py --gpus $N --steps 8 --warmup 3) > gpurun_out/r2_multi_default_n$N.log 2>&1
echo "rc=$?"; grep '^{' gpurun_out/r2_multi_default_n$N.log | cut -c1-300; grep real gpurun_out/r2_multi_default_n$N.log

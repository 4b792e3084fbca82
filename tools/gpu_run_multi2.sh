#!/bin/bash
# quick multi-GPU check of the candidate-sharded bench mode (clean shutdown) -- tight timeouts
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
(time timeout 170 $TR bench.py --gpus $N --steps 3 --warmup 3 --shard candidates --batch 8 --candidates 64 --prof-steps 1) > gpurun_out/r2_multi_cands2_n$N.log 2>&1
echo "rc=$?"; grep '^{' gpurun_out/r2_multi_cands2_n$N.log | cut -c1-400; grep real gpurun_out/r2_multi_cands2_n$N.log

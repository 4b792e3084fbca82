#!/bin/bash
# Multi-GPU call (gpurun --gpus N): the default bench line (weak + extra legs), strong scaling, candidate sharding
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/r2_multi_weak_n$N.log 2>&1
timeout 400 $TR bench.py --gpus $N --steps 6 --warmup 3 --scaling strong --prof-steps 0 > gpurun_out/r2_multi_strong_n$N.log 2>&1
timeout 400 $TR bench.py --gpus $N --steps 6 --warmup 3 --shard candidates --batch 8 --candidates 64 --prof-steps 0 > gpurun_out/r2_multi_cands_n$N.log 2>&1
timeout 400 $TR bench.py --gpus $N --steps 4 --warmup 3 --config c5 --shard candidates --prof-steps 0 > gpurun_out/r2_multi_c5_cands_n$N.log 2>&1
timeout 400 $TR bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/r2_multi_ref_n$N.log 2>&1
for f in weak strong cands c5_cands ref; do echo "== $f"; grep '^{' gpurun_out/r2_multi_${f}_n$N.log | cut -c1-1500 || tail -5 gpurun_out/r2_multi_${f}_n$N.log; grep -c . gpurun_out/r2_multi_${f}_n$N.log; done

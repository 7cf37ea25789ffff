#!/bin/bash
# 2 GPUs: sharded fused tail on both ranking routes vs the separate kernels, the multi-GPU check (tails), headline bench
N=${1:-2}
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tools/mgpu_tail_dbg.py 2>&1 | grep "^n/rank\|first diff" | tee gpurun_out/r2_tail_routes_n$N.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 tools/mgpu_greedy_check.py tails 2>&1 | grep '"tail"\|MGPU' | cut -c1-260
bash tools/gpu_n8_headline.sh $N 1

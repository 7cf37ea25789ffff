#!/bin/bash
# multi-GPU correctness of the persistent selection loop: picks vs the single-GPU loop (run with gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tools/mgpu_greedy_check.py ${2:-} > gpurun_out/mgpu_greedy_check_n$N.txt 2>&1
tail -30 gpurun_out/mgpu_greedy_check_n$N.txt

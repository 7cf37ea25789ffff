#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
ALQ_PERSIST_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    tools/mgpu_greedy_check.py bigonly > gpurun_out/mgpu_dbg_n$N.txt 2>&1
grep -v "^$" gpurun_out/mgpu_dbg_n$N.txt | grep "dbg\|picks_match" | head -120

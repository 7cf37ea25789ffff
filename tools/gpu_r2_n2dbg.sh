#!/bin/bash
# 2-GPU: headline-only bench (one rep) + phase stamps of the sharded fused tail (ALQ_SELECT_DEBUG)
N=${1:-2}
mkdir -p gpurun_out
bash tools/gpu_n8_headline.sh $N 1
ALQ_SELECT_DEBUG=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/mgpu_greedy_check.py ${2:-tailonly} > gpurun_out/r2_taildbg_n$N.log 2>&1
grep -c "fused tail dbg" gpurun_out/r2_taildbg_n$N.log
grep "rank 0\]" gpurun_out/r2_taildbg_n$N.log | tail -8
grep '"tail"' gpurun_out/r2_taildbg_n$N.log | tail -14

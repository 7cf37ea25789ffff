#!/bin/bash
# N-GPU bench exactly as the driver launches it (arg 1 = N)
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N exit=$?"; cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err

#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
for rep in $(seq 1 ${2:-2}); do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$rep bench.py --gpus $N --no-extras > gpurun_out/r2_headline_n${N}_$rep.json 2> gpurun_out/r2_headline_n${N}_$rep.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2_headline_n${N}_$rep.json'))
print(d['ms_per_step'], d['value'], d['step_kernel_ms'], d.get('picks_match_single_gpu'))
PY
done

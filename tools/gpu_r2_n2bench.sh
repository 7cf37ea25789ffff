#!/bin/bash
# 2 GPUs: the full default bench under the driver's torchrun line
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 > gpurun_out/r02c_bench_n2.json 2> gpurun_out/r02c_bench_n2.err; echo "exit=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02c_bench_n2.json'))
print(d['ms_per_step'], d['value'], d['picks_match_single_gpu'], d['e2e']['value'])
for k, w in d['workloads'].items():
    if 'us_per_selection_step' in w:
        print(' ', k, round(w['us_per_selection_step'], 2), round(w['ms_per_step'], 1), (w.get('roofline') or {}).get('frac'), w.get('picks_match_single_gpu'))
PY
tail -3 gpurun_out/r02c_bench_n2.err

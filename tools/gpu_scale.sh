#!/bin/bash
# bench.py under torchrun for N GPUs (gpurun --gpus N); prints the greedy workloads' summary
N=${1:-2}
mkdir -p gpurun_out
if [ "$2" == "pytest" ]; then timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r2_pytest_multi_n$N.txt; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -3 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_n$N.json'))
print({k:d[k] for k in ('n_gpus','value','ms_per_step','host_enqueue_ms_per_step')}, d['roofline']['frac'])
for k,v in d.get('workloads',{}).items():
    if isinstance(v,dict): print(k, {a:v.get(a) for a in ('error','value','ms_per_step','us_per_selection_step','picks_match_single_gpu','breakdown_ms')}, (v.get('roofline') or {}).get('frac'), (v.get('roofline') or {}).get('streaming_phase'))
PY

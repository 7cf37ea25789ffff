#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r2f_pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err
tail -3 gpurun_out/r2f_bench_n1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','roofline','gpu_launches')})
for k,v in d.get('workloads',{}).items():
    if isinstance(v,dict): print(k, {a:v.get(a) for a in ('error','value','ms_per_step','us_per_selection_step','loop_variant','breakdown_ms','picks_unique')}, (v.get('roofline') or {}).get('frac'))
PY

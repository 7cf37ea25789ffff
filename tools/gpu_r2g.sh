#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r2g_pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err
tail -3 gpurun_out/r2g_bench_n1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2g_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'])
for k,v in d.get('workloads',{}).items():
    if isinstance(v,dict): print(k, {a:v.get(a) for a in ('error','value','ms_per_step','us_per_selection_step','loop_variant','breakdown_ms','picks_unique')}, (v.get('roofline') or {}).get('frac'), (v.get('cpu_baseline') or {}).get('value'))
PY

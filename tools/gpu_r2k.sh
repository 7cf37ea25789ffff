#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2k_pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/r2k_bench_n1.json 2> gpurun_out/r2k_bench_n1.err
tail -3 gpurun_out/r2k_bench_n1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2k_bench_n1.json'))
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','host_enqueue_ms_per_step','latency_ms_single_query')}, d['roofline']['frac'], d['e2e']['value'], d.get('pipelined'))
print(d.get('cpu_baseline',{}).get('value'), d.get('clocks'))
for k,v in d.get('workloads',{}).items():
    if isinstance(v,dict): print(k, {a:v.get(a) for a in ('error','value','ms_per_step','us_per_selection_step','loop_variant')}, (v.get('roofline') or {}).get('frac'), (v.get('cpu_baseline') or {}).get('value'), (v.get('k3') or {}).get('frac'), (v.get('k3') or {}).get('peak'))
PY

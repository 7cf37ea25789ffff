#!/bin/bash
# 1 GPU, end of round 2: the whole gpu test tier, then the full default bench
mkdir -p gpurun_out
timeout 170 python -m pytest tests -x -q -m gpu > gpurun_out/r02c_pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/r02c_pytest_gpu.log
timeout 170 python bench.py > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err; echo "bench exit=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02c_bench_n1.json'))
print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['clocks'])
for k, v in d.items():
    if isinstance(v, dict) and ('us_per_selection_step' in v or 'ms' in v):
        print(k, {a: b for a, b in v.items() if not isinstance(b, (dict, list))})
PY

#!/bin/bash
# 1 GPU: scores + greedy parity tests, fused-tail phase stamps, headline-only bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scores.py tests/test_gpu_greedy.py -x -q -m gpu 2>&1 | tail -3
timeout 200 python tools/tail_time.py > gpurun_out/r2_tail_time.txt 2>&1; tail -5 gpurun_out/r2_tail_time.txt
timeout 300 python bench.py --no-extras 2>/dev/null | tee gpurun_out/r2_headline_n1.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['streaming_phase'], d['step_kernel_ms'], d['pipelined'], d['host_enqueue_ms_per_step'])"

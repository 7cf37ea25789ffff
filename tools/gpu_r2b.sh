#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_greedy.py -x -q -m gpu 2>&1 | tail -15
ALQ_PERSIST_DEBUG=1 PT_VARIANTS=3 timeout 300 python tools/persist_time.py > gpurun_out/r2e_persist_time.jsonl 2> gpurun_out/r2e_persist_time.err
cat gpurun_out/r2e_persist_time.jsonl; tail -8 gpurun_out/r2e_persist_time.err
ALQ_D2_FAST_PATH=0 PT_VARIANTS=3 PT_KINDS=factored timeout 300 python tools/persist_time.py 2>/dev/null | tail -1

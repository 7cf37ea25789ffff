#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_greedy.py -x -q -m gpu 2>&1 | tail -4
PT_ROWS=10000 PT_VARIANTS=3 PT_STEPS=2000 timeout 300 python tools/persist_time.py 2>/dev/null
PT_VARIANTS=3 PT_STEPS=1000 timeout 300 python tools/persist_time.py 2>/dev/null

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scores.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python tools/tail_time.py 2>&1 | tee gpurun_out/r2h_tail_time.txt

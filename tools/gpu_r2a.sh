#!/bin/bash
# round 2, first GPU call: the new persistent loop -- parity tests, then timing against the per-step launches
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_greedy.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2a_pytest_greedy.txt
cat gpurun_out/r2a_pytest_greedy.txt
timeout 300 python tools/persist_time.py > gpurun_out/r2a_persist_time.jsonl 2> gpurun_out/r2a_persist_time.err
cat gpurun_out/r2a_persist_time.jsonl; tail -5 gpurun_out/r2a_persist_time.err

#!/bin/bash
# parity tests + smoke + bench (no profiler)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err

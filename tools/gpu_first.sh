#!/bin/bash
# First GPU pass: parity tests file by file (each under its own timeout so a hung kernel cannot
# eat the lease), smoke, then a short bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for f in test_gpu_scores test_gpu_greedy test_gpu_e2e; do
  timeout 420 python -m pytest tests/$f.py -x -q -m gpu > gpurun_out/$f.log 2>&1
  echo "$f exit=$?" | tee -a gpurun_out/summary.txt
  tail -5 gpurun_out/$f.log
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err

#!/bin/bash
# K6 under ncu: gap table, the two pipelined row kernels (min-only, per-class) and BASE's candidate / resolve kernels
# as launched by tools/mase_time.py (= bench workloads.mase_base).  First iteration (5 matching launches) skipped.
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:"rows_pipe_kernel|class_gap_inv|base_candidates|base_resolve" --launch-skip 5 -c 5 \
    -f -o gpurun_out/r01f_mase python tools/mase_time.py > gpurun_out/r01f_mase.log 2>&1
echo "ncu exit=$?"
ncu -i gpurun_out/r01f_mase.ncu-rep --page raw --csv > gpurun_out/r01f_mase_raw.csv 2>/dev/null
sz=$(stat -c %s gpurun_out/r01f_mase.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 30000000 ]; then rm -f gpurun_out/r01f_mase.ncu-rep; fi
ls -la gpurun_out | grep r01f

#!/bin/bash
# K6 under ncu: the two pipelined row kernels (min-only, per-class) and the gap-table kernel of tools/mase_time.py.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mase.py -x -q 2>&1 | tail -3
timeout 120 python tools/mase_time.py 2>&1 | tail -1 > gpurun_out/r01e_mase_time.json
cat gpurun_out/r01e_mase_time.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rows_pipe_kernel|class_gap_inv" --launch-skip 3 -c 3 \
    -f -o gpurun_out/r01e_mase python tools/mase_time.py > gpurun_out/r01e_mase.log 2>&1
echo "ncu exit=$?"
ncu -i gpurun_out/r01e_mase.ncu-rep --page raw --csv > gpurun_out/r01e_mase_raw.csv 2>/dev/null
ncu -i gpurun_out/r01e_mase.ncu-rep --page details --csv > gpurun_out/r01e_mase_details.csv 2>/dev/null
sz=$(stat -c %s gpurun_out/r01e_mase.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 30000000 ]; then rm -f gpurun_out/r01e_mase.ncu-rep; fi
ls -la gpurun_out | grep r01e

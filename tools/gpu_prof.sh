#!/bin/bash
# ncu passes (B200_PROFILING.md): launch list with device times, then --set full on the hot kernels.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/prof_target.py > gpurun_out/prof_launch.log 2>&1
echo "launch list exit=$?"
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:"score_rows_vec|step_pipe|step_direct|sample_kernel|min_dist_kernel|select_|sort_single|badge_factors" \
    -f -o gpurun_out/prof_full python tools/prof_target.py > gpurun_out/prof_full.log 2>&1
echo "full exit=$?"
ls -la gpurun_out/

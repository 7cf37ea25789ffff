#!/bin/bash
# r01d: launch list of the final build + full captures of the hot kernels (one ncu invocation per workload)
mkdir -p gpurun_out
export PROF_STEPS=4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_final.csv python tools/prof_target.py > gpurun_out/prof_launch.log 2>&1
echo "launch list exit=$?"
full() {
  name=$1; rx=$2; cnt=$3; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" -c $cnt \
      -f -o gpurun_out/$name python tools/prof_target.py "$@" > gpurun_out/$name.log 2>&1
  echo "$name exit=$?"
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  sz=$(stat -c %s gpurun_out/$name.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 9000000 ]; then rm -f gpurun_out/$name.ncu-rep; fi
}
full prof_d_margin  "rows_pipe_kernel|select_cluster_kernel" 6 margin
full prof_d_coreset "min_dist_tc_kernel|step_pipe_kernel" 3 coreset
full prof_d_badge   "min_dist_tc_kernel|step_pipe_kernel|sample_cluster_kernel|rows_pipe_kernel" 8 badge
du -sh gpurun_out

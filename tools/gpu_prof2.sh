#!/bin/bash
# r01b: launch list of the current build + full captures of K3 (tensor), K1 pipe, sampling cluster
mkdir -p gpurun_out
export PROF_STEPS=4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_b.csv python tools/prof_target.py > gpurun_out/prof_launch.log 2>&1
echo "launch list exit=$?"
full() {
  name=$1; rx=$2; cnt=$3; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" -c $cnt \
      -f -o gpurun_out/$name python tools/prof_target.py "$@" > gpurun_out/$name.log 2>&1
  echo "$name exit=$?"
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  sz=$(stat -c %s gpurun_out/$name.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 12000000 ]; then rm -f gpurun_out/$name.ncu-rep; fi
}
full prof_b_k3      "min_dist_tc_kernel|split_tf32" 3 coreset
full prof_b_sample  "sample_cluster_kernel|step_pipe" 5 badge
full prof_b_margin  "rows_pipe_kernel|select_|sort_runs|merge_rank" 10 margin
du -sh gpurun_out

#!/bin/bash
# ncu passes (B200_PROFILING.md): launch list with device times, then --set full on the hot kernels.
# Reports are exported to CSV on the box; .ncu-rep files above 15 MB are dropped (gpurun_out <= 64 MiB).
mkdir -p gpurun_out
export PROF_STEPS=4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/prof_target.py > gpurun_out/prof_launch.log 2>&1
echo "launch list exit=$?"
full() {  # name, kernel regex, count, targets...
  name=$1; rx=$2; cnt=$3; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" -c $cnt \
      -f -o gpurun_out/$name python tools/prof_target.py "$@" > gpurun_out/$name.log 2>&1
  echo "$name exit=$?"
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  sz=$(stat -c %s gpurun_out/$name.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 15000000 ]; then
     ncu -i gpurun_out/$name.ncu-rep --page source --csv > gpurun_out/${name}_source.csv 2>/dev/null
     rm -f gpurun_out/$name.ncu-rep
  fi
}
full prof_margin  "score_rows_vec|select_|sort_single" 14 margin
full prof_coreset "step_pipe|step_direct" 4 coreset
full prof_badge   "step_pipe|step_direct|sample_kernel" 8 badge
du -sh gpurun_out; ls -la gpurun_out/

#!/bin/bash
# round 2 ncu evidence: launch list of the bench command + --set full captures of the three hot kernels
mkdir -p gpurun_out
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --extra-steps 1 > gpurun_out/r02_launches_bench.log 2>&1
tail -2 gpurun_out/r02_launches_bench.log | cut -c1-300
for t in margin coreset badge; do
  case $t in margin) K="regex:rows_pipe_kernel";; *) K="regex:greedy_persist_kernel";; esac
  PROF_STEPS=24 timeout 600 ncu --set full --clock-control none --import-source on -k $K -s 1 -c 1 -o gpurun_out/r02_ncu_$t -f \
      python tools/prof_target.py $t > gpurun_out/r02_ncu_$t.log 2>&1
  tail -1 gpurun_out/r02_ncu_$t.log
  ncu -i gpurun_out/r02_ncu_$t.ncu-rep --page raw --csv > gpurun_out/r02_ncu_${t}_raw.csv 2>/dev/null
  ls -la gpurun_out/r02_ncu_$t.ncu-rep | awk '{print $5}'
done
PROF_STEPS=24 timeout 600 ncu --set full --clock-control none -k regex:min_dist_tc -c 1 -o gpurun_out/r02_ncu_k3 -f python tools/prof_target.py k3 > gpurun_out/r02_ncu_k3.log 2>&1
ncu -i gpurun_out/r02_ncu_k3.ncu-rep --page raw --csv > gpurun_out/r02_ncu_k3_raw.csv 2>/dev/null
# keep the reports small enough to travel
for f in gpurun_out/r02_ncu_*.ncu-rep; do s=$(stat -c %s $f); if [ $s -gt 25000000 ]; then rm $f; fi; done
ls -la gpurun_out | tail -15

#!/bin/bash
# usage: gpu_prof_one.sh <name> <kernel-regex> <count> <prof_target args...>
name=$1; rx=$2; cnt=$3; shift 3
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" -c $cnt \
    -f -o gpurun_out/$name python tools/prof_target.py "$@" > gpurun_out/$name.log 2>&1
echo "$name exit=$?"
ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
ncu -i gpurun_out/$name.ncu-rep --page details --csv > gpurun_out/${name}_details.csv 2>/dev/null
sz=$(stat -c %s gpurun_out/$name.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 20000000 ]; then rm -f gpurun_out/$name.ncu-rep; fi
ls -la gpurun_out | head -20

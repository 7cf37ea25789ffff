#!/bin/bash
N=${1:-8}
bash tools/gpu_multi.sh $N big | tail -16
bash tools/gpu_scale.sh $N

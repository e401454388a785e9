#!/bin/bash
# ncu --set full of one kernel family ($1 = regex, default k_assign) in batch mode (64 pairs); report -> gpurun_out/prof_$2.ncu-rep
mkdir -p gpurun_out
K=${1:-k_assign}; N=${2:-a}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 4 -c 3 \
   -f -o gpurun_out/prof_$N python tools/prof_batch.py 64 > gpurun_out/prof_$N.log 2>&1
tail -n 4 gpurun_out/prof_$N.log

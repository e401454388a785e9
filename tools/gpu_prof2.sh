#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_assign|k_search|k_reduce_solve|k_colscan' -s 24 -c 8 \
   -f -o gpurun_out/prof_batch2 python tools/prof_batch.py 64 > gpurun_out/prof_batch2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_assign|k_search|k_reduce_solve' -s 9 -c 6 \
   -f -o gpurun_out/prof_single2 python tools/prof_target.py single 4 > gpurun_out/prof_single2.log 2>&1
tail -n 4 gpurun_out/prof_batch2.log gpurun_out/prof_single2.log

#!/bin/bash
# ncu --set full captures of the fused kernels: batch mode (64 pairs) and latency mode (1 pair).
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_assign|k_search|k_reduce_solve|k_colscan' -s 24 -c 8 \
   -f -o gpurun_out/prof_batch python tools/prof_batch.py 64 > gpurun_out/prof_batch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_assign|k_search|k_reduce_solve|k_colscan' -s 11 -c 8 \
   -f -o gpurun_out/prof_single python tools/prof_target.py single 4 > gpurun_out/prof_single.log 2>&1
python tools/latency_breakdown.py > gpurun_out/latency_breakdown.log 2>&1
tail -5 gpurun_out/prof_batch.log gpurun_out/prof_single.log; cat gpurun_out/latency_breakdown.log

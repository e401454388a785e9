#!/bin/bash
# Round-2 measurement pass on one B200 (everything profiles/r02_* is made from).  Output -> gpurun_out/
mkdir -p gpurun_out
cp profiles/ncu_summary.json gpurun_out/ncu_summary.json
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_smi.txt 2>&1
( time timeout 600 python bench.py ) > gpurun_out/r02_bench.log 2> gpurun_out/r02_bench.err
( time timeout 300 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r02_bench_ref.log 2>&1
# launch list of the bench command (stream launches: kernels that use the device-side graph API are not listed when replayed from a graph)
ICP_B200_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches.csv \
   python bench.py --steps 1 --warmup 3 --pairs 256 --no-cpu-baseline --no-scaled --pairs-total 0 > gpurun_out/r02_bench_under_ncu.log 2>&1
# ncu --set full: the four batch kernels, unfused, at iteration 10 of a 64-pair registration
ICP_B200_FUSED=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_assign|k_search|k_reduce_solve|k_colscan' -s 42 -c 4 \
   -f -o gpurun_out/r02_full_batch64 python tools/prof_batch2.py 64 14 > gpurun_out/r02_full_batch64.log 2>&1
# ... and of one pair in latency mode (4th iteration)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_assign|k_search|k_reduce_solve|k_colscan' -s 14 -c 4 \
   -f -o gpurun_out/r02_full_single python tools/prof_target.py single 6 > gpurun_out/r02_full_single.log 2>&1
# ... and of one registration at BASELINE config 4's largest size (307200 landmarks / 1024 representatives), 10th iteration
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_assign|k_search|k_reduce|k_colscan' -s 65 -c 7 \
   -f -o gpurun_out/r02_scaled_full python tools/prof_scaled.py > gpurun_out/r02_scaled_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_scaled_launches.csv \
   python tools/prof_scaled.py > gpurun_out/r02_scaled_launches.log 2>&1
( timeout 200 python tools/cprime_phases.py; timeout 200 python tools/a_phases.py; timeout 200 python tools/d_phases.py ) > gpurun_out/r02_batch_phase_clocks.log 2>&1
# summaries are made on the box (gpurun copies back at most 64 MiB: the three reports together exceed it; the single-pair report is dropped)
python tools/ncu_report.py gpurun_out/r02_full_batch64.ncu-rep "round 2, batch mode, 64 pairs per launch, iteration 10 of a registration, kernels launched unfused in stream order" --json gpurun_out/ncu_summary.json --key-suffix _batch --pairs 64 > gpurun_out/r02_ncu_full_batch64.md 2>&1
python tools/ncu_report.py gpurun_out/r02_full_single.ncu-rep "round 2, one pair in latency mode, 4th iteration" --json gpurun_out/ncu_summary.json --key-suffix _single --pairs 1 > gpurun_out/r02_ncu_full_single.md 2>&1
python tools/ncu_report.py gpurun_out/r02_scaled_full.ncu-rep "round 2, one registration of 307200 points / 1024 representatives (BASELINE config 4), 10th iteration" > gpurun_out/r02_ncu_full_scaled_307200_1024.md 2>&1
rm -f gpurun_out/r02_full_single.ncu-rep gpurun_out/r02_scaled_full.ncu-rep      # only the batch report (~27 MB) travels back
timeout 300 python tools/latency_breakdown.py > gpurun_out/r02_latency_breakdown.log 2>&1
timeout 300 python tools/latency_engines.py > gpurun_out/r02_latency_engines.log 2>&1
timeout 300 python tools/scaled_ab.py > gpurun_out/r02_scaled.log 2>&1
tail -c 300 gpurun_out/r02_bench.err; tail -2 gpurun_out/r02_bench_ref.log | cut -c1-300; tail -3 gpurun_out/r02_full_batch64.log; tail -3 gpurun_out/r02_full_single.log; cat gpurun_out/r02_latency_engines.log | cut -c1-400; cat gpurun_out/r02_scaled.log; tail -3 gpurun_out/r02_scaled_full.log; cat gpurun_out/r02_batch_phase_clocks.log

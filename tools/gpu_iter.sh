#!/bin/bash
# quick iteration pass: parity tests, then knob sweeps ($1 = A|C|L|all|none), then latency breakdown
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 5 gpurun_out/pytest_gpu.log
if [ "${1:-all}" != "none" ]; then
  for w in ${1//,/ }; do timeout 1200 python tools/tune2.py $w; done > gpurun_out/tune2.log 2>&1
  cat gpurun_out/tune2.log
fi
timeout 300 python tools/latency_breakdown.py > gpurun_out/latency_breakdown.log 2>&1
cat gpurun_out/latency_breakdown.log

#!/bin/bash
# parity tests (optional: $1 = notest to skip), then a short list of tune configs given as env-style strings in $2..
mkdir -p gpurun_out
if [ "$1" != "notest" ]; then
  ( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
  tail -n 4 gpurun_out/pytest_gpu.log
fi
shift
for cfg in "$@"; do
  echo "== $cfg"; env $cfg python -c "import sys; sys.path.insert(0,'tools'); from tune import CHILD; exec(CHILD)"
done

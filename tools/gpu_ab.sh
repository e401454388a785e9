#!/bin/bash
# parity tests on the default build, then the kernel-A knob sweep on the default build and on an alternative build ($1)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 5 gpurun_out/pytest_gpu.log
echo "== default build"; timeout 600 python tools/tune2.py A
if [ -n "$1" ]; then echo "== $1"; ICP_B200_LIB=$PWD/$1 timeout 600 python tools/tune2.py A; fi

#!/bin/bash
# A/B builds: tools/build_variant.sh NAME "-DMACRO=.. ..." -> build/NAME/libicp_b200.so (icp_fused.cu recompiled with the
# extra flags, the other objects reused from icp_b200/csrc).  Use with ICP_B200_LIB=build/NAME/libicp_b200.so.
set -e
cd "$(dirname "$0")/../icp_b200/csrc"
make -s
mkdir -p ../../build/$1
NVFLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC -cudart static -Xptxas -v"
/usr/local/cuda/bin/nvcc $NVFLAGS $2 -c icp_fused.cu -o ../../build/$1/icp_fused.o 2> ../../build/$1/icp_fused.ptxas.log || (cat ../../build/$1/icp_fused.ptxas.log; false)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../../build/$1/libicp_b200.so \
   icp_api.o icp_stages.o icp_engine.o ../../build/$1/icp_fused.o icp_batch.o icp_multi.o icp_exact.o icp_bench.o -Xlinker --version-script=exports.map
grep -A2 "k_search_grouped" ../../build/$1/icp_fused.ptxas.log | grep Used

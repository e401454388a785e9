#!/bin/bash
# kernel C A/B: default build and the build/ variants at several QG (queries per CTA); per-kernel ms at TUNE_PAIRS pairs
export TUNE_PAIRS=${TUNE_PAIRS:-256} ICP_B200_BATCH_SLICES=1
run() { echo "== $*"; env "$@" python -c "import sys; sys.path.insert(0,'tools'); from tune import CHILD; exec(CHILD)"; }
run X=default
run ICP_B200_QG=512
run ICP_B200_QG=2048
for v in "$@"; do
  IFS=: read name qgs <<< "$v"
  for qg in ${qgs//,/ }; do run ICP_B200_LIB=$PWD/build/$name/libicp_b200.so ICP_B200_QG=$qg; done
done

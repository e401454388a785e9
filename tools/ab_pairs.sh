#!/bin/bash
# pairs-per-step / slices sweep of the batched bench (wave quantisation of the 8-CTA-per-pair kernels): usage tools/ab_pairs.sh "PAIRS SLICES" ...
for ps in "$@"; do
  set -- $ps
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-scaled --pairs-total 0 --pairs $1 --slices $2 > gpurun_out/abp.log 2> gpurun_out/abp.err || tail -3 gpurun_out/abp.err
  python - "$ps" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/abp.log").read().strip().splitlines()[-1])
    print(f"pairs/slices {sys.argv[1]:10s} value {d['value']:9.1f} e2e {d['e2e']['value']:9.1f} us/pair-iter {d['us_per_pair_iteration_batched']:.3f}")
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done

#!/bin/bash
# A/B of kernel-C flavours on the benchmarked workload: one short bench.py run per environment setting.
# usage: tools/ab_bench.sh "NAME=VALUE ..." "NAME=VALUE ..." ...   ("-" = default environment)
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  i=$((i+1))
  if [ "$envs" = "-" ]; then e=""; else e="$envs"; fi
  env $e python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-scaled --pairs-total 0 > gpurun_out/ab_$i.log 2> gpurun_out/ab_$i.err || tail -5 gpurun_out/ab_$i.err
  python - "$envs" gpurun_out/ab_$i.log <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1]); r = d["roofline"]
    ms = {k: round(v, 4) for k, v in r["kernel_ms_per_launch"].items()}
    print(f"{sys.argv[1]:40s} value {d['value']:9.1f} e2e {d['e2e']['value']:9.1f} ms/step {d['ms_per_step']:.3f} {ms} frac {r['frac']:.3f}")
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done

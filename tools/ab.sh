#!/bin/bash
# A/B timing of env-switched variants on one GPU box:  tools/ab.sh TAG "ENV1=a ENV2=b" "ENV1=c" ...
# every variant runs bench.py --skip-extras twice (interleaved) and prints ms_per_step + the phase split
tag=$1; shift
for rep in 1 2; do
  i=0
  for v in "$@"; do
    i=$((i+1))
    env $v python bench.py --skip-extras --no-cpu-baseline --steps 400 > gpurun_out/${tag}_v${i}_r${rep}.json 2> gpurun_out/${tag}_v${i}_r${rep}.err
    python - "$v" gpurun_out/${tag}_v${i}_r${rep}.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[2]) if l.startswith("{")][-1])
    print(sys.argv[1], "ms_per_step %.4f" % d["ms_per_step"], {k: round(v * 1e3, 1) for k, v in d["phases_ms"].items()})
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  done
done

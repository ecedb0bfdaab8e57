#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/r2c18_summary.txt
: > $S
one() {  # label, args
  timeout 300 python bench.py --lean --steps 300 --warmup 10 --no-strong $2 > gpurun_out/r2c18_tmp.json 2>/dev/null
  python - <<PY | tee -a $S
import json
try:
    a = json.loads(open("gpurun_out/r2c18_tmp.json").read().strip().splitlines()[-1])
    print("$1: apply us", round(a["ms_per_step"] * 1e3, 2), "launches", a["launches_per_step"])
except Exception as e:
    print("$1: failed", e)
PY
}
for cfg in 0 1 2 3 4; do one "cfg$cfg auto" "--config $cfg"; done
for v in 400 600 800 1000; do one "cfg1 head $v/1000" "--config 1 --variant $v"; done
for v in 150 250 350 450; do one "cfg4 head $v/1000" "--config 4 --variant $v"; done
for v in 400 500 600; do one "cfg2 head $v/1000" "--config 2 --variant $v"; done

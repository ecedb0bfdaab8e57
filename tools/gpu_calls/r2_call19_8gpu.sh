#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/r2c19_summary.txt
: > $S
one() {  # label, nproc, args
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $2 --lean --steps 300 --warmup 10 --no-strong $3 > gpurun_out/r2c19_tmp.json 2>/dev/null
  python - <<PY | tee -a $S
import json
try:
    a = json.loads(open("gpurun_out/r2c19_tmp.json").read().strip().splitlines()[-1])
    print("$1: apply us", round(a["ms_per_step"] * 1e3, 2), "value", round(a["value"]), "launches", a["launches_per_step"], "parity", (a.get("parity") or {}).get("ok"))
except Exception as e:
    print("$1: failed", e)
PY
}
one "N=8 auto" 8 ""
for v in 350 550 650; do one "N=8 head $v/1000" 8 "--variant $v"; done
one "N=4 auto" 4 ""
summ() {
python - <<PY | tee -a $S
import json
try:
    a = json.loads(open("gpurun_out/r2c19_$1.json").read().strip().splitlines()[-1])
    s4 = a.get("strong_scaling_config4") or {}
    print("$1", "value", round(a["value"], 1), "us", round(a["ms_per_step"] * 1e3, 2), "e2e", round(a["e2e"]["value"], 1), "h2d", a["e2e"]["h2d_bytes_per_step"],
          "parity", a.get("parity"), "setup", round(a["setup_ms"], 2), "strong4", round(s4.get("applies_per_s", 0), 1), s4.get("error"))
except Exception as e:
    print("$1 failed", e)
PY
}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 8 --steps 200 --warmup 10 > gpurun_out/r2c19_n8.json 2> gpurun_out/r2c19_n8.err; summ n8

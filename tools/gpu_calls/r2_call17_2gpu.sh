#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/r2c17_summary.txt
: > $S
timeout 300 python -m pytest tests/test_gpu_zz_limits.py tests/test_zzz_gpu_sharded.py -m gpu -q -x -k "graph_and_plain or peer_memory or nccl" > gpurun_out/r2c17_tests.log 2>&1
echo "tests rc=$?" | tee -a $S; tail -2 gpurun_out/r2c17_tests.log | tee -a $S
one() {  # label, nproc, args
  if [ "$2" = "1" ]; then
    timeout 300 python bench.py --lean --steps 300 --warmup 10 --no-strong $3 > gpurun_out/r2c17_tmp.json 2>/dev/null
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $2 --lean --steps 300 --warmup 10 --no-strong $3 > gpurun_out/r2c17_tmp.json 2>/dev/null
  fi
  python - <<PY | tee -a $S
import json
try:
    a = json.loads(open("gpurun_out/r2c17_tmp.json").read().strip().splitlines()[-1])
    print("$1: apply us", round(a["ms_per_step"] * 1e3, 2), "value", round(a["value"]), "launches", a["launches_per_step"], "parity", (a.get("parity") or {}).get("ok"))
except Exception as e:
    print("$1: failed", e)
PY
}
one "N=1 auto" 1 ""
for v in 200 270 350 450; do one "N=1 head $v/1000" 1 "--variant $v"; done
one "N=2 auto" 2 ""
for v in 200 270 350 450 550; do one "N=2 head $v/1000" 2 "--variant $v"; done

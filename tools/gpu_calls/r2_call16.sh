#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/r2c16_summary.txt
: > $S
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_limits.py tests/test_gpu_pcg.py tests/test_zzz_gpu_sharded.py -m gpu -q -x > gpurun_out/r2c16_tests.log 2>&1
echo "tests rc=$?" | tee -a $S
grep -E "^FAILED|passed|failed|^E  " gpurun_out/r2c16_tests.log | head -12 | tee -a $S
one() {  # label, args
  timeout 200 python bench.py --lean --steps 300 --warmup 10 $2 > gpurun_out/r2c16_tmp.json 2>/dev/null
  python - <<PY | tee -a $S
import json
try:
    a = json.loads(open("gpurun_out/r2c16_tmp.json").read().strip().splitlines()[-1])
    print("$1: apply us", round(a["ms_per_step"] * 1e3, 2), "launches", a["launches_per_step"])
except Exception as e:
    print("$1: failed", e)
PY
}
for cfg in 0 1 2; do
  one "cfg$cfg separate kernels" "--config $cfg --fused-chain 0"
  one "cfg$cfg fused chain auto head" "--config $cfg"
done
for v in 100 150 200 250 300 400; do one "cfg2 fused head $v/1000" "--config 2 --variant $v"; done
for v in 300 500 700 1000; do one "cfg1 fused head $v/1000" "--config 1 --variant $v"; done

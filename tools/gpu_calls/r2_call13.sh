#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/r2c13_summary.txt
: > $S
timeout 300 python tools/invert_variant_bench.py 1024 1,0 2>&1 | tail -1 | tee -a $S
MAS_B200_LIB=$PWD/preconditioner-for-cloth-and-deformable-body-simulation_b200/libmas_b200_phase.so MAS_PHASE_TIMING=1 \
  timeout 120 python tools/invert_variant_bench.py 1024 0 2>&1 | grep -m1 "phase cycles" | tee -a $S
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2c13_gpu_tests.log 2>&1
echo "gpu suite rc=$?" | tee -a $S
grep -E "^FAILED|passed|failed" gpurun_out/r2c13_gpu_tests.log | tail -5 | tee -a $S
timeout 900 python bench.py --config 1 --proximity --no-strong > gpurun_out/r2c13_cfg1prox.json 2> gpurun_out/r2c13_cfg1prox.err
python - <<'PY' | tee -a $S
import json
try:
    a = json.loads(open("gpurun_out/r2c13_cfg1prox.json").read().strip().splitlines()[-1])
    print("cfg1prox:", a["config"]["workload"], "value", round(a["value"], 1), "setup", a["setup_device_ms"], "parity", a.get("parity"), "cpu", a.get("cpu_baseline"), "pcg", a.get("pcg"))
except Exception as e:
    print("cfg1prox failed", e)
PY
tail -3 gpurun_out/r2c13_cfg1prox.err | cut -c1-300 | tee -a $S

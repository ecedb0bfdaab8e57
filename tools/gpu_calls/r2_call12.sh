#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/r2c12_summary.txt
: > $S
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2c12_gpu_tests.log 2>&1
echo "gpu suite rc=$?" | tee -a $S
grep -E "^FAILED|passed|failed" gpurun_out/r2c12_gpu_tests.log | tail -10 | tee -a $S
grep -E "^E  " gpurun_out/r2c12_gpu_tests.log | head -10 | tee -a $S
summ() {
python - <<PY | tee -a $S
import json
try:
    a = json.loads(open("gpurun_out/r2c12_$1.json").read().strip().splitlines()[-1])
    print("$1:", "value", round(a["value"], 1), "us", round(a["ms_per_step"] * 1e3, 2), "setup", a["setup_device_ms"], "rebuild", a.get("setup_rebuild_hierarchy_ms"),
          "e2e", round(a["e2e"]["value"], 1), "whole-apply frac", a["roofline"]["whole_apply"]["frac"] if a.get("roofline") else None,
          "pcg", (a["pcg"]["iterations"], round(a["pcg"]["solve_ms"], 2)) if a.get("pcg") else None, "parity", a.get("parity"), "strong4", a.get("strong_scaling_config4"),
          "cpu", (round(a["cpu_baseline"]["value"], 1), round(a["cpu_baseline"]["setup_ms"], 1)) if a.get("cpu_baseline") else None)
except Exception as e:
    print("$1 failed", e)
PY
}
timeout 600 python bench.py > gpurun_out/r2c12_cfg2.json 2> gpurun_out/r2c12_cfg2.err; summ cfg2
for c in 0 1 3 4; do
  timeout 900 python bench.py --config $c --no-strong > gpurun_out/r2c12_cfg$c.json 2> gpurun_out/r2c12_cfg$c.err; summ cfg$c
done
timeout 600 python bench.py --config 1 --proximity --no-strong > gpurun_out/r2c12_cfg1prox.json 2> gpurun_out/r2c12_cfg1prox.err; summ cfg1prox
tail -3 gpurun_out/r2c12_cfg1prox.err | cut -c1-300 | tee -a $S

#!/bin/bash
# Last single-GPU run of the round (after the five-level cross_bank fix): whole GPU suite, smoke, default bench, configs 3 and 4.
mkdir -p gpurun_out/final_d
O=gpurun_out/final_d
timeout 1500 python -m pytest tests -m gpu -q > $O/gpu_tests_1gpu.log 2>&1; echo "gpu suite rc=$?"
grep -E "^FAILED|passed|failed" $O/gpu_tests_1gpu.log | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > $O/bench_1gpu_b.json 2> $O/bench_1gpu_b.err; echo "bench rc=$?"
for c in 3 4; do timeout 900 python bench.py --config $c --no-strong > $O/bench_cfg$c.json 2> $O/bench_cfg$c.err; done
python - <<'PY'
import json
for name in ("bench_1gpu_b", "bench_cfg3", "bench_cfg4"):
    a = json.loads(open(f"gpurun_out/final_d/{name}.json").read().strip().splitlines()[-1])
    print(name, "value", round(a["value"], 1), "us", round(a["ms_per_step"] * 1e3, 2), "setup", round(a["setup_device_ms"], 3), "rebuild", round(a.get("setup_rebuild_hierarchy_ms", 0), 3), "e2e", round(a["e2e"]["value"], 1),
          "frac", round(a["roofline"]["whole_apply"]["frac"], 4), "setup roofline", round(a["roofline"]["setup"]["ms"], 3), "pcg", a["pcg"]["iterations"], round(a["pcg"]["solve_ms"], 2), a["pcg"].get("us_per_iteration_steady"), a["parity"].get("ok"), a["clocks"]["sm_mhz"], a["clocks"]["reasons"])
PY

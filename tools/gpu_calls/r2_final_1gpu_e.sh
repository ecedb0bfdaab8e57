#!/bin/bash
# Very last single-GPU run of the round (after the collision-Hessian pre-summation): whole GPU suite, smoke, config 1 in both
# stencil flavours, setup timelines.
mkdir -p gpurun_out/final_e
O=gpurun_out/final_e
timeout 1500 python -m pytest tests -m gpu -q > $O/gpu_tests_1gpu.log 2>&1; echo "gpu suite rc=$?"
grep -E "^FAILED|passed|failed" $O/gpu_tests_1gpu.log | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --config 1 --no-strong > $O/bench_cfg1.json 2> $O/bench_cfg1.err
timeout 900 python bench.py --config 1 --proximity --no-strong > $O/bench_cfg1_proximity_stencils.json 2> $O/bench_cfg1_proximity.err
python - <<'PY'
import json
for name in ("bench_cfg1", "bench_cfg1_proximity_stencils"):
    a = json.loads(open(f"gpurun_out/final_e/{name}.json").read().strip().splitlines()[-1])
    print(name, "value", round(a["value"], 1), "us", round(a["ms_per_step"] * 1e3, 2), "setup", round(a["setup_device_ms"], 3), "rebuild", round(a.get("setup_rebuild_hierarchy_ms", 0), 3),
          "pcg", a["pcg"]["iterations"], round(a["pcg"]["solve_ms"], 2), a["parity"].get("ok"), a["parity"].get("rel_l2_gpu_vs_f64"), a["parity"].get("rel_l2_reference_vs_f64"), "cpu", round(a["cpu_baseline"]["value"], 1), a["cpu_baseline"]["kind"])
PY
MAS_CONFIG=1 MAS_PROXIMITY=1 timeout 600 python tools/profile_setup.py > $O/setup_timeline_cfg1_proximity.txt 2>/dev/null
MAS_CONFIG=1 timeout 600 python tools/profile_setup.py > $O/setup_timeline_cfg1.txt 2>/dev/null
grep -E "prepare device|sum by" $O/setup_timeline_cfg1_proximity.txt $O/setup_timeline_cfg1.txt

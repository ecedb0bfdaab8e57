#!/bin/bash
# Round 2, last single-GPU evidence run on the final tree: whole GPU suite, smoke, bench (both arms, every config), ncu launch
# lists of the bench and of the PCG iteration, ncu --set full of the dominant apply kernel (DRAM traffic per launch).
mkdir -p gpurun_out
O=gpurun_out/final_c
mkdir -p $O
S=$O/summary.txt
: > $S
timeout 1500 python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "gpu suite rc=$?" | tee -a $S
grep -E "^FAILED|passed|failed" $O/gpu_tests.log | tail -5 | tee -a $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee -a $S
timeout 900 python bench.py > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench rc=$?" | tee -a $S
timeout 600 python bench.py --impl reference > $O/bench_1gpu_reference_arm.json 2>/dev/null; echo "reference arm rc=$?" | tee -a $S
for c in 0 1 3 4; do timeout 900 python bench.py --config $c --no-strong > $O/bench_cfg$c.json 2> $O/bench_cfg$c.err; done
timeout 900 python bench.py --config 1 --proximity --no-strong > $O/bench_cfg1_proximity_stencils.json 2> $O/bench_cfg1_proximity.err
python - <<'PY' | tee -a $S
import json
for name in ("bench_1gpu", "bench_cfg0", "bench_cfg1", "bench_cfg1_proximity_stencils", "bench_cfg3", "bench_cfg4"):
    try:
        a = json.loads(open(f"gpurun_out/final_c/{name}.json").read().strip().splitlines()[-1])
        par = a.get("parity") or {}
        pcg = a.get("pcg") or {}
        print(name, "value", round(a["value"], 1), "us", round(a["ms_per_step"] * 1e3, 2), "setup", round(a["setup_device_ms"], 3), "rebuild", round(a.get("setup_rebuild_hierarchy_ms", 0), 3),
              "e2e", round(a["e2e"]["value"], 1), a["e2e"].get("applies_per_s_by_staging"), "frac", round(a["roofline"]["frac"], 4), round(a["roofline"]["whole_apply"]["frac"], 4),
              "pcg", pcg.get("iterations"), round(pcg.get("solve_ms", 0), 2), "par", par.get("rel_l2_gpu_vs_f64"), par.get("rel_l2_reference_vs_f64"), par.get("ok"),
              "cpu", round(a["cpu_baseline"]["value"], 1), round(a["cpu_baseline"].get("setup_ms", 0), 1), "strong4", (a.get("strong_scaling_config4") or {}).get("applies_per_s"))
    except Exception as e:
        print(name, "failed", e)
try:
    a = json.loads(open("gpurun_out/final_c/bench_1gpu_reference_arm.json").read().strip().splitlines()[-1])
    print("reference arm", round(a["value"], 2), a["steps"], a["warmup"], a["cpu_baseline"]["cores"])
except Exception as e:
    print("reference arm failed", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_lean.csv python bench.py --lean --steps 20 --warmup 3 --no-strong > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?" | tee -a $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_pcg_iteration.csv python tools/pcg_kernels.py 6 > $O/ncu_launches_pcg.log 2>&1; echo "ncu pcg launches rc=$?" | tee -a $S
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_fine -s 40 -c 3 -f -o $O/solve_fine python bench.py --lean --steps 20 --warmup 3 --no-strong > $O/ncu_solve_fine.log 2>&1; echo "ncu solve_fine rc=$?" | tee -a $S
ls -la $O | tail -24

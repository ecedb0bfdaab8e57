#!/bin/bash
# Round 2, final single-GPU evidence run: probe, whole GPU suite, smoke, bench (both arms, all configs), ncu launch list and
# ncu --set full of the tensor-core inversion, per-phase cycles, maximum-size check.
mkdir -p gpurun_out
O=gpurun_out/final
mkdir -p $O
S=$O/summary.txt
: > $S
timeout 120 ./tools/tcgen05_probe > $O/tcgen05_probe.txt 2>&1; echo "probe rc=$?" | tee -a $S
timeout 1500 python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "gpu suite rc=$?" | tee -a $S
grep -E "^FAILED|passed|failed" $O/gpu_tests.log | tail -5 | tee -a $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee -a $S
timeout 900 python bench.py > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench rc=$?" | tee -a $S
timeout 600 python bench.py --impl reference > $O/bench_1gpu_reference_arm.json 2>/dev/null; echo "reference arm rc=$?" | tee -a $S
for c in 0 1 3 4; do timeout 900 python bench.py --config $c --no-strong > $O/bench_cfg$c.json 2> $O/bench_cfg$c.err; done
python - <<'PY' | tee -a $S
import json
for name in ("bench_1gpu", "bench_cfg0", "bench_cfg1", "bench_cfg3", "bench_cfg4"):
    try:
        a = json.loads(open(f"gpurun_out/final/{name}.json").read().strip().splitlines()[-1])
        par = a.get("parity") or {}
        print(name, "value", round(a["value"], 1), "us", round(a["ms_per_step"] * 1e3, 2), "setup", round(a["setup_device_ms"], 3), "rebuild", round(a["setup_rebuild_hierarchy_ms"], 3),
              "e2e", round(a["e2e"]["value"], 1), a["e2e"]["applies_per_s_by_staging"], "frac", round(a["roofline"]["frac"], 4), round(a["roofline"]["whole_apply"]["frac"], 4),
              "pcg", a["pcg"]["iterations"], round(a["pcg"]["solve_ms"], 2), "par", par.get("rel_l2_gpu_vs_f64"), par.get("rel_l2_reference_vs_f64"), par.get("ok"),
              "cpu", round(a["cpu_baseline"]["value"], 1), round(a["cpu_baseline"]["setup_ms"], 1), "strong4", (a.get("strong_scaling_config4") or {}).get("applies_per_s"))
    except Exception as e:
        print(name, "failed", e)
try:
    a = json.loads(open("gpurun_out/final/bench_1gpu_reference_arm.json").read().strip().splitlines()[-1])
    print("reference arm", round(a["value"], 2), a["steps"], a["warmup"], a["cpu_baseline"]["cores"])
except Exception as e:
    print("reference arm failed", e)
PY
MAS_B200_LIB=$PWD/preconditioner-for-cloth-and-deformable-body-simulation_b200/libmas_b200_phase.so MAS_PHASE_TIMING=1 \
  timeout 120 python tools/invert_variant_bench.py 1024 0 2>&1 | grep -m1 "phase cycles" > $O/invert_phase_cycles.txt
timeout 300 python tools/invert_variant_bench.py 1024 1,0 2>&1 | tail -1 > $O/invert_variants.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_lean.csv python bench.py --lean --steps 20 --warmup 3 --no-strong > $O/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fine_assemble_invert_tc -s 2 -c 1 -o $O/invert_tc -f python tools/invert_variant_bench.py 1024 0 > $O/ncu_invert.log 2>&1
timeout 600 python tools/max_size_check.py > $O/max_size_33m_verts.json 2> $O/max_size.err; tail -c 400 $O/max_size_33m_verts.json | tee -a $S
ls -la $O | tail -20

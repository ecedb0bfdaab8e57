#!/bin/bash
# Round 2, second final single-GPU run (after the PCG harness work): whole GPU suite, smoke, bench on every config, launch list.
mkdir -p gpurun_out
O=gpurun_out/final_b
mkdir -p $O
S=$O/summary.txt
: > $S
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
        a = json.loads(open(f"gpurun_out/final_b/{name}.json").read().strip().splitlines()[-1])
        par = a.get("parity") or {}
        print(name, "value", round(a["value"], 1), "us", round(a["ms_per_step"] * 1e3, 2), "setup", round(a["setup_device_ms"], 3),
              "e2e", round(a["e2e"]["value"], 1), "frac", round(a["roofline"]["frac"], 4), round(a["roofline"]["whole_apply"]["frac"], 4),
              "pcg", a["pcg"]["iterations"], round(a["pcg"]["solve_ms"], 2), round(1e3 * a["pcg"]["ms_per_iteration"], 1), "par", par.get("rel_l2_gpu_vs_f64"), par.get("rel_l2_reference_vs_f64"), par.get("ok"),
              "cpu", round(a["cpu_baseline"]["value"], 1))
    except Exception as e:
        print(name, "failed", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_pcg_iteration.csv python tools/pcg_kernels.py 6 > $O/ncu_launches.log 2>&1; echo "ncu rc=$?" | tee -a $S

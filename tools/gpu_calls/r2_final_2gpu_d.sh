#!/bin/bash
# Last check of the round on two GPUs: whole GPU suite (sharded tests included), default bench line on one GPU, weak line on two.
mkdir -p gpurun_out/final_d
O=gpurun_out/final_d
timeout 1500 python -m pytest tests -m gpu -q > $O/gpu_tests_2gpu.log 2>&1; echo "gpu suite (2 GPUs visible) rc=$?"
grep -E "^FAILED|passed|failed" $O/gpu_tests_2gpu.log | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > $O/bench_1gpu_reference_arm.json 2>/dev/null; echo "reference arm rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29573 bench.py --gpus 2 --steps 200 --warmup 10 > $O/bench_2gpu_weak.json 2> $O/bench_2gpu_weak.err; echo "n2 rc=$?"
python - <<'PY'
import json
a = json.loads(open("gpurun_out/final_d/bench_1gpu.json").read().strip().splitlines()[-1])
print("N=1", round(a["value"], 1), round(a["ms_per_step"] * 1e3, 2), "setup", round(a["setup_device_ms"], 3), "e2e", round(a["e2e"]["value"], 1), "frac", round(a["roofline"]["frac"], 4), round(a["roofline"]["whole_apply"]["frac"], 4), "traffic", a["roofline"]["traffic"], {k: v for k, v in a["pcg"].items() if k != "timing"}, a["parity"].get("ok"), a["clocks"], "cpu", round(a["cpu_baseline"]["value"], 1))
b = json.loads(open("gpurun_out/final_d/bench_1gpu_reference_arm.json").read().strip().splitlines()[-1])
print("reference arm", round(b["value"], 2), b["steps"], b["warmup"], b["config"] == a["config"])
c = json.loads(open("gpurun_out/final_d/bench_2gpu_weak.json").read().strip().splitlines()[-1])
print("N=2", round(c["value"], 1), round(c["ms_per_step"] * 1e3, 2), round(c["e2e"]["value"], 1), c["parity"], c["clocks"], (c.get("strong_scaling_config4") or {}).get("applies_per_s"))
PY

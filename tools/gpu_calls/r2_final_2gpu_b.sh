#!/bin/bash
# Round 2, final multi-GPU check on two GPUs: the whole GPU suite (the sharded tests need two devices), then the weak bench.
mkdir -p gpurun_out/final_b
O=gpurun_out/final_b
timeout 1500 python -m pytest tests -m gpu -q > $O/gpu_tests_2gpu.log 2>&1; echo "gpu suite (2 GPUs visible) rc=$?"
grep -E "^FAILED|passed|failed" $O/gpu_tests_2gpu.log | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 200 --warmup 10 > $O/bench_2gpu_weak.json 2> $O/bench_2gpu_weak.err
echo "n2 rc=$?"
python - <<'PY'
import json
a = json.loads(open("gpurun_out/final_b/bench_2gpu_weak.json").read().strip().splitlines()[-1])
print("N=2", round(a["value"], 1), round(a["ms_per_step"] * 1e3, 2), round(a["e2e"]["value"], 1), a["parity"], (a.get("strong_scaling_config4") or {}).get("applies_per_s"))
PY

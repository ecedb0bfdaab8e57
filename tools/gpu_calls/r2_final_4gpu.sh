#!/bin/bash
mkdir -p gpurun_out/final
for n in 2 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2957$n bench.py --gpus $n --steps 200 --warmup 10 > gpurun_out/final/bench_${n}gpu_weak.json 2> gpurun_out/final/bench_${n}gpu_weak.err
  echo "n$n rc=$?"
  python - <<PY
import json
a = json.loads(open("gpurun_out/final/bench_${n}gpu_weak.json").read().strip().splitlines()[-1])
print("N=$n", round(a["value"], 1), round(a["ms_per_step"] * 1e3, 2), round(a["e2e"]["value"], 1), a["parity"], (a.get("strong_scaling_config4") or {}).get("applies_per_s"))
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29579 bench.py --gpus 4 --impl reference --steps 20 --warmup 3 > gpurun_out/final/bench_4gpu_reference_arm.json 2>/dev/null; tail -c 300 gpurun_out/final/bench_4gpu_reference_arm.json

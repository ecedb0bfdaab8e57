#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 2 --steps 200 --warmup 10 --no-strong > gpurun_out/r2_clock_2gpu.json 2> gpurun_out/r2_clock_2gpu.err; echo "n2 rc=$?"
timeout 600 python bench.py --config 0 --no-strong --no-arbiter > gpurun_out/r2_clock_cfg0.json 2>/dev/null; echo "cfg0 rc=$?"
python - <<'PY'
import json
for f in ("r2_clock_2gpu", "r2_clock_cfg0"):
    a = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, a["n_gpus"], round(a["value"], 1), round(a["ms_per_step"] * 1e3, 2), a["clocks"], (a.get("parity") or {}).get("ok"))
PY

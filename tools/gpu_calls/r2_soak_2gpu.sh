#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zz_limits.py -x -q -k "three_thousand" > gpurun_out/r2_soak_tests.log 2>&1; echo "soak tests rc=$?"; tail -3 gpurun_out/r2_soak_tests.log
# 20,000 sharded applies in a row through the peer-memory exchange (apply counter, flag parity, double-buffered send slots),
# then the bench's own check of the merged z against a single GPU
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 2 --steps 20000 --warmup 10 --no-strong > gpurun_out/r2_soak_2gpu.json 2> gpurun_out/r2_soak_2gpu.err; echo "soak bench rc=$?"
python - <<'PY'
import json
a = json.loads(open("gpurun_out/r2_soak_2gpu.json").read().strip().splitlines()[-1])
print("N=2", a["steps"], round(a["value"], 1), round(a["ms_per_step"] * 1e3, 2), a["parity"], a["clocks"])
PY

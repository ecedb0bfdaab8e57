#!/bin/bash
# 2-GPU box: the sharded suite with real NCCL ranks, then the bench at N = 2 (weak + strong) with in-run parity.
mkdir -p gpurun_out
S=gpurun_out/r2c5_summary.txt
: > $S
nvidia-smi --query-gpu=index,name --format=csv,noheader | tee -a $S
timeout 600 python -m pytest tests/test_zzz_gpu_sharded.py -m gpu -q > gpurun_out/r2c5_sharded.log 2>&1
echo "sharded rc=$?" | tee -a $S
grep -E "^FAILED|^E  |passed|failed|skipped" gpurun_out/r2c5_sharded.log | head -20 | tee -a $S
run() {  # name, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 $2 > gpurun_out/r2c5_$1.json 2> gpurun_out/r2c5_$1.err
  echo "$1 rc=$?" | tee -a $S
  python - <<PY | tee -a $S
import json
try:
    a = json.loads(open("gpurun_out/r2c5_$1.json").read().strip().splitlines()[-1])
    print("$1", "value", round(a["value"], 1), "ms", round(a["ms_per_step"], 4), "e2e", round(a["e2e"]["value"], 1), a["e2e"]["h2d_bytes_per_step"], "parity", a["parity"], "setup", a["setup_ms"])
except Exception as e:
    print("$1 failed", e)
PY
  tail -2 gpurun_out/r2c5_$1.err | cut -c1-300 | tee -a $S
}
run weak ""
run strong_cfg4 "--config 4"
run weak_nccl "--nccl-exchange"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 20 --warmup 3 > gpurun_out/r2c5_ref.json 2>/dev/null
python -c "
import json; a=json.loads(open('gpurun_out/r2c5_ref.json').read().strip().splitlines()[-1]); print('reference arm', round(a['value'],1), a['config']['workload'][:80])" | tee -a $S

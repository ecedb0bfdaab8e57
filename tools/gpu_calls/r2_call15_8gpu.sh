#!/bin/bash
# 8-GPU box: the driver's SCALE sequence (default command at N = 2, 4, 8; N = 1 for the same box) with in-run parity,
# plus BASELINE config 4 strong scaling inside the default line (strong_scaling_config4).
mkdir -p gpurun_out
S=gpurun_out/r2c15_summary.txt
: > $S
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 | tee -a $S
summ() {
python - <<PY | tee -a $S
import json
try:
    a = json.loads(open("gpurun_out/r2c15_$1.json").read().strip().splitlines()[-1])
    s4 = a.get("strong_scaling_config4") or {}
    print("$1", "value", round(a["value"], 1), "us", round(a["ms_per_step"] * 1e3, 2), "e2e", round(a["e2e"]["value"], 1), "h2d", a["e2e"]["h2d_bytes_per_step"],
          "parity", a.get("parity"), "setup", round(a["setup_ms"], 2), "strong4", round(s4.get("applies_per_s", 0), 1), s4.get("error"),
          "fine_ms", (a.get("roofline") or {}).get("per_rank_fine_kernel_ms"))
except Exception as e:
    print("$1 failed", e)
PY
}
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c15_n1.json 2> gpurun_out/r2c15_n1.err; summ n1
for n in 2 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 200 --warmup 10 > gpurun_out/r2c15_n$n.json 2> gpurun_out/r2c15_n$n.err
  echo "n$n rc=$?" | tee -a $S
  summ n$n
  grep -E "Error|error|Traceback" gpurun_out/r2c15_n$n.err | head -3 | cut -c1-300 | tee -a $S
done
timeout 600 python -m pytest tests/test_zzz_gpu_sharded.py -m gpu -q -k "nccl" > gpurun_out/r2c15_nccl_test.log 2>&1
tail -2 gpurun_out/r2c15_nccl_test.log | tee -a $S

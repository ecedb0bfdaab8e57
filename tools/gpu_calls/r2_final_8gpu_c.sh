#!/bin/bash
# Round 2, final tree on an 8-GPU box: weak-scaling bench lines at 8 and 4 GPUs (in-run parity) and the reference arm under torchrun.
mkdir -p gpurun_out/final_c
O=gpurun_out/final_c
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2958$n bench.py --gpus $n --steps 200 --warmup 10 > $O/bench_${n}gpu_weak.json 2> $O/bench_${n}gpu_weak.err
  echo "n$n rc=$?"
  python - <<PY
import json
a = json.loads(open("gpurun_out/final_c/bench_${n}gpu_weak.json").read().strip().splitlines()[-1])
print("N=$n", round(a["value"], 1), round(a["ms_per_step"] * 1e3, 2), round(a["e2e"]["value"], 1), a["parity"], (a.get("strong_scaling_config4") or {}).get("applies_per_s"), a["clocks"])
PY
done
timeout 300 python bench.py --lean --steps 200 --warmup 10 --no-strong > $O/bench_1gpu_on_8gpu_box.json 2>/dev/null; python -c "
import json
a = json.loads(open('gpurun_out/final_c/bench_1gpu_on_8gpu_box.json').read().strip().splitlines()[-1]); print('N=1 (lean)', round(a['value'], 1), round(a['ms_per_step'] * 1e3, 2))"

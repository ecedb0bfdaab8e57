#!/bin/bash
# Round 2, GPU call 2: bring-up of the tcgen05 inversion kernel (probe of the primitives first), then the suites.
mkdir -p gpurun_out
S=gpurun_out/r2c2_summary.txt
: > $S
timeout 120 ./tools/tcgen05_probe 2>&1 | tee -a $S
echo "probe rc=$?" | tee -a $S
timeout 300 python tools/invert_variant_bench.py 256 1,0 2>&1 | tail -2 | tee -a $S
timeout 300 python tools/invert_variant_bench.py 1024 1,0 2>&1 | tail -2 | tee -a $S
MAS_B200_LIB=$PWD/preconditioner-for-cloth-and-deformable-body-simulation_b200/libmas_b200_phase.so MAS_PHASE_TIMING=1 \
  timeout 120 python tools/invert_variant_bench.py 1024 0 2>&1 | grep -m2 "phase cycles" | tee -a $S
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c2_gpu_tests.log 2>&1
echo "gpu suite rc=$?" | tee -a $S
tail -5 gpurun_out/r2c2_gpu_tests.log | tee -a $S
timeout 400 python bench.py > gpurun_out/r2c2_bench_1gpu.json 2> gpurun_out/r2c2_bench_1gpu.err
python - <<'PY' | tee -a $S
import json
try:
    a = json.loads(open("gpurun_out/r2c2_bench_1gpu.json").read().strip().splitlines()[-1])
    print({k: a[k] for k in ("value", "ms_per_step", "setup_ms", "setup_device_ms")}, a["e2e"]["value"], a["roofline"]["frac"], a["pcg"])
except Exception as e:
    print("bench failed", e)
PY
tail -3 gpurun_out/r2c2_bench_1gpu.err | tee -a $S

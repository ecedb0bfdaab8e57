#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/r2c4_summary.txt
: > $S
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_limits.py -m gpu -q -k "full_size_values or cached_hierarchy" > gpurun_out/r2c4_tests.log 2>&1
echo "tests rc=$?" | tee -a $S
grep -E "^FAILED|^E  |passed|failed" gpurun_out/r2c4_tests.log | head -30 | tee -a $S
timeout 600 python bench.py > gpurun_out/r2c4_bench_1gpu.json 2> gpurun_out/r2c4_bench_1gpu.err
echo "bench rc=$?" | tee -a $S
tail -5 gpurun_out/r2c4_bench_1gpu.err | tee -a $S
python - <<'PY' | tee -a $S
import json
try:
    a = json.loads(open("gpurun_out/r2c4_bench_1gpu.json").read().strip().splitlines()[-1])
    for k in ("value", "ms_per_step", "setup_ms", "setup_device_ms", "setup_rebuild_hierarchy_ms", "e2e", "parity", "pcg", "clocks"):
        print(k, a.get(k))
    print("roofline", {k: v for k, v in a["roofline"].items() if k in ("achieved", "frac", "setup")})
except Exception as e:
    print("bench parse failed", e)
PY
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2c4_bench_ref.json 2>/dev/null
python -c "
import json; a=json.loads(open('gpurun_out/r2c4_bench_ref.json').read().strip().splitlines()[-1]); print('reference arm', a['value'], a['steps'], a['warmup'], sorted(a['config']))" | tee -a $S

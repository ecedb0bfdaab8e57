#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_small_suite.log 2>&1; echo "gpu suite rc=$?"; grep -E "^FAILED|passed|failed" gpurun_out/r2_small_suite.log | tail -5
for c in 0 1; do timeout 600 python bench.py --config $c --no-strong --steps 400 --warmup 20 > gpurun_out/r2_small_cfg$c.json 2> gpurun_out/r2_small_cfg$c.err; python -c "
import json
a = json.loads(open('gpurun_out/r2_small_cfg$c.json').read().strip().splitlines()[-1]); print('cfg$c', round(a['value'], 1), round(a['ms_per_step'] * 1e3, 2), a['gpu_launches'], a['parity'].get('ok'), a['pcg']['iterations'], round(a['pcg']['solve_ms'], 2))"; done
MAS_N=64 timeout 300 python tools/profile_sharded.py > gpurun_out/r2_small_timeline64.txt 2>/dev/null; grep -E "kernel  |span|next" gpurun_out/r2_small_timeline64.txt

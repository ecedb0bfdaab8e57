#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pcg.py -x -q > gpurun_out/r2_pcg_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_pcg_tests.log
timeout 300 python tools/profile_pcg.py 32 > gpurun_out/r2_pcg_timeline.txt 2> gpurun_out/r2_pcg_timeline.err
echo "rc=$?"; cat gpurun_out/r2_pcg_timeline.txt | head -40; tail -3 gpurun_out/r2_pcg_timeline.err
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "full_size" > gpurun_out/r2_pcg_tests2.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_pcg_tests2.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-strong --no-arbiter > gpurun_out/r2_pcg_bench.json 2> gpurun_out/r2_pcg_bench.err; echo "bench rc=$?"
python -c "
import json
a = json.loads(open('gpurun_out/r2_pcg_bench.json').read().strip().splitlines()[-1]); print(a['pcg'])"

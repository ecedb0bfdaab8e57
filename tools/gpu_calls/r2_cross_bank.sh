#!/bin/bash
mkdir -p gpurun_out
for n in 1024 2048; do MAS_N=$n timeout 300 python tools/profile_setup.py > gpurun_out/r2_setup_timeline_${n}_after.txt 2>/dev/null; cat gpurun_out/r2_setup_timeline_${n}_after.txt; done
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_limits.py -x -q -k "full_size or five_level or maximum_size or cached_hierarchy" > gpurun_out/r2_cross_bank_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2_cross_bank_tests.log
timeout 600 python tools/max_size_check.py 2>/dev/null | tail -c 600

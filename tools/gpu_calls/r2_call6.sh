#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/r2c6_summary.txt
: > $S
timeout 300 python tools/invert_variant_bench.py 1024 1,0 2>&1 | tail -2 | tee -a $S
MAS_B200_LIB=$PWD/preconditioner-for-cloth-and-deformable-body-simulation_b200/libmas_b200_phase.so MAS_PHASE_TIMING=1 \
  timeout 120 python tools/invert_variant_bench.py 1024 0 2>&1 | grep -m1 "phase cycles" | tee -a $S
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2c6_gpu_tests.log 2>&1
echo "gpu suite rc=$?" | tee -a $S
grep -E "^FAILED|passed|failed" gpurun_out/r2c6_gpu_tests.log | tail -30 | tee -a $S
grep -E "^E  " gpurun_out/r2c6_gpu_tests.log | head -20 | tee -a $S

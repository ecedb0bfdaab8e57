#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/r2c14_summary.txt
: > $S
timeout 300 python tools/invert_variant_bench.py 1024 1,0 2>&1 | tail -1 | tee -a $S
MAS_B200_LIB=$PWD/preconditioner-for-cloth-and-deformable-body-simulation_b200/libmas_b200_phase.so MAS_PHASE_TIMING=1 \
  timeout 120 python tools/invert_variant_bench.py 1024 0 2>&1 | grep -m1 "phase cycles" | tee -a $S
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2c14_tests.log 2>&1
echo "parity rc=$?" | tee -a $S
tail -2 gpurun_out/r2c14_tests.log | tee -a $S

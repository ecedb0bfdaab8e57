#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/r2c9_summary.txt
: > $S
for m in 0 1; do
  MAS_TC_PIVOT=$m timeout 300 python tools/invert_variant_bench.py 1024 1,0 2>&1 | tail -1 | sed "s/^/pivot $m: /" | tee -a $S
  MAS_TC_PIVOT=$m MAS_B200_LIB=$PWD/preconditioner-for-cloth-and-deformable-body-simulation_b200/libmas_b200_phase.so MAS_PHASE_TIMING=1 \
    timeout 120 python tools/invert_variant_bench.py 1024 0 2>&1 | grep -m1 "phase cycles" | sed "s/^/pivot $m: /" | tee -a $S
done

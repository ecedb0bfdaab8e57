#!/bin/bash
mkdir -p gpurun_out
for n in 1024 2048; do MAS_N=$n timeout 300 python tools/profile_setup.py > gpurun_out/r2_setup_timeline_$n.txt 2>/dev/null; cat gpurun_out/r2_setup_timeline_$n.txt; done

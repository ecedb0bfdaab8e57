#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fine_assemble_invert_tc -s 2 -c 1 \
    -o gpurun_out/r2c7_invert_tc -f python tools/invert_variant_bench.py 1024 0 > gpurun_out/r2c7_ncu.log 2>&1
tail -3 gpurun_out/r2c7_ncu.log
ls -la gpurun_out/r2c7_invert_tc.ncu-rep

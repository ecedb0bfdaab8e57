#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_pcg.py 32 > gpurun_out/r2_pcg_timeline.txt 2> gpurun_out/r2_pcg_timeline.err
echo "rc=$?"; cat gpurun_out/r2_pcg_timeline.txt | head -40
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "full_size" > gpurun_out/r2_pcg_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_pcg_tests.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_dot -s 20 -c 1 -f -o gpurun_out/r2_spmv python tools/pcg_kernels.py 30 > gpurun_out/r2_spmv_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2_spmv_ncu.log

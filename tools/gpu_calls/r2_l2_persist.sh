#!/bin/bash
mkdir -p gpurun_out
for v in 0 1; do
  MAS_PCG_PERSIST_L2=$v timeout 300 python tools/profile_pcg.py 32 > gpurun_out/r2_persist${v}_pcg.txt 2> gpurun_out/r2_persist${v}_pcg.err
  echo "== persist=$v pcg rc=$?"; grep -E "kernel  |median|iterations" gpurun_out/r2_persist${v}_pcg.txt; tail -2 gpurun_out/r2_persist${v}_pcg.err | cut -c1-200
done
for v in 0 1; do
  MAS_N=2048 MAS_PCG_PERSIST_L2=$v timeout 300 python tools/profile_pcg.py 16 > gpurun_out/r2_persist${v}_pcg2048.txt 2> gpurun_out/r2_persist${v}_pcg2048.err
  echo "== persist=$v pcg 2048 rc=$?"; grep -E "median|iterations" gpurun_out/r2_persist${v}_pcg2048.txt
  MAS_N=512 MAS_PCG_PERSIST_L2=$v timeout 300 python tools/profile_pcg.py 32 > gpurun_out/r2_persist${v}_pcg512.txt 2> gpurun_out/r2_persist${v}_pcg512.err
  echo "== persist=$v pcg 512 rc=$?"; grep -E "median|iterations" gpurun_out/r2_persist${v}_pcg512.txt
done
timeout 900 python -m pytest tests/test_gpu_pcg.py -x -q > gpurun_out/r2_persist_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_persist_tests.log

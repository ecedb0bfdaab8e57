#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zz_limits.py -x -q -k "two_contexts or lifecycle or three_thousand or stencil_counts" > gpurun_out/r2_threads.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2_threads.log

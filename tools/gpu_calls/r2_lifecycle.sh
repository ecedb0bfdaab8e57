#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zz_limits.py -x -q -k "lifecycle" > gpurun_out/r2_lifecycle.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2_lifecycle.log

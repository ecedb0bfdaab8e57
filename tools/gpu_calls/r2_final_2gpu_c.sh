#!/bin/bash
mkdir -p gpurun_out/final_c
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/final_c/gpu_tests_2gpu.log 2>&1; echo "gpu suite (2 GPUs visible) rc=$?"
grep -E "^FAILED|passed|failed" gpurun_out/final_c/gpu_tests_2gpu.log | tail -5

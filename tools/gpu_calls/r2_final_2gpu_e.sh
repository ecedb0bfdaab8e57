#!/bin/bash
mkdir -p gpurun_out/final_e
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/final_e/gpu_tests_2gpu.log 2>&1; echo "gpu suite (2 GPUs visible) rc=$?"
grep -E "^FAILED|passed|failed" gpurun_out/final_e/gpu_tests_2gpu.log | tail -5

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pcg.py tests/test_gpu_zz_limits.py tests/test_cabi.py -x -q -k "collision or proximity or config1 or cpp_caller or stencil or lifecycle or cached" > gpurun_out/r2_collision_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2_collision_tests.log
MAS_CONFIG=1 MAS_PROXIMITY=1 timeout 600 python tools/profile_setup.py 2>/dev/null | grep -E "prepare device|collision_hessian|sum by"
MAS_CONFIG=1 timeout 600 python tools/profile_setup.py 2>/dev/null | grep -E "prepare device|collision_hessian|sum by"

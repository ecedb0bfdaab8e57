#!/bin/bash
mkdir -p gpurun_out
MAS_N=1024 timeout 300 python tools/profile_setup.py 2>/dev/null | grep -E "prepare device|cross_bank"
MAS_CONFIG=3 timeout 300 python tools/profile_setup.py 2>/dev/null | grep -E "prepare device|cross_bank"
MAS_N=2048 timeout 300 python tools/profile_setup.py 2>/dev/null | grep -E "prepare device|cross_bank"

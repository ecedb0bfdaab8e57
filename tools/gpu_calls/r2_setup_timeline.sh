#!/bin/bash
mkdir -p gpurun_out
MAS_CONFIG=1 MAS_PROXIMITY=1 timeout 600 python tools/profile_setup.py > gpurun_out/r2_setup_timeline_cfg1p.txt 2>gpurun_out/r2_setup_timeline_cfg1p.err; cat gpurun_out/r2_setup_timeline_cfg1p.txt; tail -3 gpurun_out/r2_setup_timeline_cfg1p.err

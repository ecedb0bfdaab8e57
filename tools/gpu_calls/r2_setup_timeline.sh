#!/bin/bash
mkdir -p gpurun_out
for c in 1 3; do MAS_CONFIG=$c timeout 300 python tools/profile_setup.py > gpurun_out/r2_setup_timeline_cfg$c.txt 2>/dev/null; cat gpurun_out/r2_setup_timeline_cfg$c.txt; done
timeout 600 python tools/max_size_check.py > gpurun_out/max_size_33m_verts.json 2> gpurun_out/max_size.err; tail -c 300 gpurun_out/max_size_33m_verts.json

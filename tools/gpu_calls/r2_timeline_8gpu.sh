#!/bin/bash
mkdir -p gpurun_out
MAS_NX=4096 MAS_NY=2048 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 tools/profile_sharded.py > gpurun_out/r2_timeline_8gpu.txt 2> gpurun_out/r2_timeline_8gpu.err
echo "rc=$?"; grep -E "^----|start|span|next" gpurun_out/r2_timeline_8gpu.txt | head -40
timeout 300 python tools/profile_sharded.py > gpurun_out/r2_timeline_1gpu.txt 2> gpurun_out/r2_timeline_1gpu.err
grep -E "^----|start|span|next" gpurun_out/r2_timeline_1gpu.txt | head -20
rm -f gpurun_out/trace_w*.json

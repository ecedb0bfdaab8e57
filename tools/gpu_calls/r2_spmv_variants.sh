#!/bin/bash
mkdir -p gpurun_out
for v in interleaved blocked batch3 b3o5 b2o6 b2o8; do
  MAS_B200_LIB=$PWD/build_tmp/libmas_$v.so timeout 300 python tools/profile_pcg.py 32 > gpurun_out/r2_spmv_$v.txt 2> gpurun_out/r2_spmv_$v.err
  echo "$v rc=$?"; grep -E "spmv_dot|median|iterations" gpurun_out/r2_spmv_$v.txt
done

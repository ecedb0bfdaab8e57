#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_pcg.py 32 > gpurun_out/r2_l2_spmvhint_pcg.txt 2> gpurun_out/r2_l2_spmvhint_pcg.err
echo "== spmv-hint-only pcg rc=$?"; grep -E "kernel  |median|iterations" gpurun_out/r2_l2_spmvhint_pcg.txt
MAS_N=2048 timeout 300 python tools/profile_pcg.py 16 > gpurun_out/r2_l2_spmvhint_pcg2048.txt 2> gpurun_out/r2_l2_spmvhint_pcg2048.err
echo "== spmv-hint-only pcg 2048 rc=$?"; grep -E "spmv|median|iterations" gpurun_out/r2_l2_spmvhint_pcg2048.txt
MAS_B200_LIB=$PWD/build_tmp/libmas_nohint.so MAS_N=2048 timeout 300 python tools/profile_pcg.py 16 > gpurun_out/r2_l2_nohint_pcg2048.txt 2> gpurun_out/r2_l2_nohint_pcg2048.err
echo "== nohint pcg 2048 rc=$?"; grep -E "spmv|median|iterations" gpurun_out/r2_l2_nohint_pcg2048.txt

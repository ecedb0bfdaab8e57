#!/bin/bash
# Round 2, GPU call 1: measure everything round 1 wrote without a GPU, plus the fixed sharded tests.
mkdir -p gpurun_out
S=gpurun_out/r2c1_summary.txt
: > $S
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | tee -a $S
# 1. inversion variants: setup time on the 1M cloth
timeout 180 python tools/invert_variant_bench.py > gpurun_out/r2c1_invert_variants.json 2> gpurun_out/r2c1_invert_variants.err
tail -1 gpurun_out/r2c1_invert_variants.json | tee -a $S
# 1b. per-phase cycles of the inversion kernel, shipped and variants 1, 4 (development build)
for v in 0 1 4; do
  MAS_B200_LIB=$PWD/preconditioner-for-cloth-and-deformable-body-simulation_b200/libmas_b200_phase.so MAS_PHASE_TIMING=1 \
    timeout 120 python tools/invert_variant_bench.py 1024 $v 2>&1 | grep -m2 "phase cycles" | sed "s/^/variant $v: /" | tee -a $S
done
# 2. experimental tests (gated)
MAS_EXPERIMENTAL=1 timeout 500 python -m pytest tests/test_gpu_zz_limits.py -m gpu -q -rA -k "experimental or register_host or cached_hierarchy or apply_chain" > gpurun_out/r2c1_experimental_tests.log 2>&1
echo "experimental tests rc=$?" | tee -a $S
grep -E "passed|failed" gpurun_out/r2c1_experimental_tests.log | tail -2 | tee -a $S
# 3. the sharded suite first (the round-1 failure), then the whole GPU suite
timeout 600 python -m pytest tests/test_zzz_gpu_sharded.py -m gpu -q -x > gpurun_out/r2c1_sharded.log 2>&1
echo "sharded rc=$?" | tee -a $S
tail -3 gpurun_out/r2c1_sharded.log | tee -a $S
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c1_gpu_tests.log 2>&1
echo "gpu suite rc=$?" | tee -a $S
tail -3 gpurun_out/r2c1_gpu_tests.log | tee -a $S
# 4. bench line + apply-chain variants
timeout 400 python bench.py > gpurun_out/r2c1_bench_1gpu.json 2> gpurun_out/r2c1_bench_1gpu.err
tail -c 1500 gpurun_out/r2c1_bench_1gpu.json | tee -a $S
for cfg in 0 1 2; do
  for chain in 0 1 7; do
    timeout 200 python bench.py --config $cfg --lean --steps 300 --warmup 10 --apply-chain $chain > gpurun_out/r2c1_chain${chain}_cfg$cfg.json 2>/dev/null
    python - <<PY | tee -a $S
import json
try:
    a = json.loads(open("gpurun_out/r2c1_chain${chain}_cfg$cfg.json").read().strip().splitlines()[-1])
    print("config $cfg chain $chain: apply us", round(a["ms_per_step"] * 1e3, 2))
except Exception as e:
    print("config $cfg chain $chain: failed", e)
PY
  done
done
# 5. ncu --set full of the inversion kernel, shipped / variant 1 / variant 4 (512^2: 8,192 blocks per launch)
for v in 0 1 4; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:fine_assemble_invert -s 2 -c 1 \
    -o gpurun_out/r2c1_invert_v$v -f python tools/invert_variant_bench.py 512 $v > gpurun_out/r2c1_ncu_invert_v$v.log 2>&1
done
ls -la gpurun_out | tail -20

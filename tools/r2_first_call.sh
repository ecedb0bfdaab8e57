#!/bin/bash
# First GPU call of a new round: everything that was written without a GPU gets measured in one go.
#   gpurun --timeout 900 -- 'bash tools/r2_first_call.sh'
# Results land in gpurun_out/r2_*.  Order: cheap and decisive first.
mkdir -p gpurun_out
# 1. the experimental inversion variants: parity (gated tests) and setup time against the shipped kernel
MAS_EXPERIMENTAL=1 timeout 400 python -m pytest tests/test_gpu_zz_limits.py -m gpu -q -rA > gpurun_out/r2_experimental_tests.log 2>&1
echo "experimental tests rc=$?" | tee -a gpurun_out/r2_summary.txt
timeout 120 python tools/invert_variant_bench.py > gpurun_out/r2_invert_variants.json 2> gpurun_out/r2_invert_variants.err
tail -1 gpurun_out/r2_invert_variants.json | tee -a gpurun_out/r2_summary.txt
# 1b. ncu --set full of the inversion kernel, shipped and experimental (512^2: 8,192 blocks per launch; one launch each)
for v in 0 1 4; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:fine_assemble_invert -s 2 -c 1 \
    -o gpurun_out/r2_invert_v$v -f python tools/invert_variant_bench.py 512 $v > gpurun_out/r2_ncu_invert_v$v.log 2>&1
done
# 2. the bench line (now with both host stagings in e2e) and the reference arm
timeout 400 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
tail -c 600 gpurun_out/r2_bench_1gpu.json | tee -a gpurun_out/r2_summary.txt
timeout 200 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_reference.json 2>/dev/null
for cfg in 0 1 2; do
  timeout 200 python bench.py --config $cfg --lean --steps 200 --warmup 10 > gpurun_out/r2_chain0_cfg$cfg.json 2>/dev/null
  timeout 200 python bench.py --config $cfg --lean --steps 200 --warmup 10 --apply-chain 7 > gpurun_out/r2_chain1_cfg$cfg.json 2>/dev/null
  python - <<PY | tee -a gpurun_out/r2_summary.txt
import json
a = json.loads(open("gpurun_out/r2_chain0_cfg$cfg.json").read().strip().splitlines()[-1]); b = json.loads(open("gpurun_out/r2_chain1_cfg$cfg.json").read().strip().splitlines()[-1])
print("config $cfg: apply us shipped", round(a["ms_per_step"] * 1e3, 2), " MAS_OPT_APPLY_CHAIN", round(b["ms_per_step"] * 1e3, 2))
PY
done
# 3. PCIe staging on this box
timeout 60 python tools/pcie_bound.py > gpurun_out/r2_pcie_bound.json 2>/dev/null
# 4. the whole GPU suite
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_gpu_tests.log 2>&1
echo "gpu suite rc=$?" | tee -a gpurun_out/r2_summary.txt
tail -3 gpurun_out/r2_gpu_tests.log | tee -a gpurun_out/r2_summary.txt

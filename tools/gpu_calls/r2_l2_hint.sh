#!/bin/bash
mkdir -p gpurun_out
for v in nohint hint; do
  if [ $v = nohint ]; then export MAS_B200_LIB=$PWD/build_tmp/libmas_nohint.so; else unset MAS_B200_LIB; fi
  timeout 300 python tools/profile_pcg.py 32 > gpurun_out/r2_l2_${v}_pcg.txt 2> gpurun_out/r2_l2_${v}_pcg.err
  echo "== $v pcg rc=$?"; grep -E "kernel  |median|iterations" gpurun_out/r2_l2_${v}_pcg.txt
  timeout 300 python tools/profile_sharded.py > gpurun_out/r2_l2_${v}_apply.txt 2> gpurun_out/r2_l2_${v}_apply.err
  echo "== $v apply rc=$?"; grep -E "kernel  |span|next" gpurun_out/r2_l2_${v}_apply.txt
  MAS_N=512 timeout 300 python tools/profile_sharded.py > gpurun_out/r2_l2_${v}_apply512.txt 2> gpurun_out/r2_l2_${v}_apply512.err
  echo "== $v apply 512 rc=$?"; grep -E "kernel  |span|next" gpurun_out/r2_l2_${v}_apply512.txt
done
unset MAS_B200_LIB
timeout 600 python bench.py --steps 200 --warmup 10 --no-strong > gpurun_out/r2_l2_hint_bench.json 2> gpurun_out/r2_l2_hint_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
a = json.loads(open("gpurun_out/r2_l2_hint_bench.json").read().strip().splitlines()[-1])
print(round(a["value"], 1), round(a["ms_per_step"] * 1e3, 2), a["roofline"]["frac"], a["roofline"]["whole_apply"]["frac"], a["parity"], a.get("pcg"))
PY
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pcg.py -x -q > gpurun_out/r2_l2_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_l2_tests.log

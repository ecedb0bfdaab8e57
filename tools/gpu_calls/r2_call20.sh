#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/r2c20_summary.txt
: > $S
cat > /tmp/san.py <<'PY'
import importlib, numpy as np, sys
sys.path.insert(0, ".")
pkg = importlib.import_module("preconditioner-for-cloth-and-deformable-body-simulation_b200")
S = pkg.synth
m = S.cloth(40, with_topology=True)
m = S.add_collisions(m, 60, 60, 120)
g = pkg.SeSchwarzPreconditioner(0).setup_from_mesh(m)
r = S.residual(m.nv)
z = np.zeros_like(r)
g.Preconditioning(z, r)
print("z norm", float(np.linalg.norm(z)))
PY
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san.py > gpurun_out/r2c20_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a $S
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|z norm|Error|hazard" gpurun_out/r2c20_$tool.log | head -8 | tee -a $S
done

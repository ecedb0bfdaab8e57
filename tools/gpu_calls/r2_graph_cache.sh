#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_graph_cache_suite.log 2>&1; echo "gpu suite rc=$?"; grep -E "^FAILED|passed|failed" gpurun_out/r2_graph_cache_suite.log | tail -5
python - <<'PY'
import importlib, os, sys, time
import torch
sys.path.insert(0, os.getcwd())
pkg = importlib.import_module("preconditioner-for-cloth-and-deformable-body-simulation_b200")
S = pkg.synth
mesh = S.cloth_rect_device(1024, 1024, torch.device("cuda:0"))
g = pkg.SeSchwarzPreconditioner(0)
g.m_positions, g.m_neighbours = mesh.positions, (mesh.nbr_starts, mesh.nbr_idx)
g.AllocatePrecoditioner(mesh.nv, 0, 0)
g.PreparePreconditioner(mesh.diag, mesh.offdiag, mesh.nbr_starts)
r = torch.from_numpy(S.residual(mesh.nv)).cuda()
for pairs in (1, 2, 4, 6):
    zs = [torch.empty_like(r) for _ in range(pairs)]
    for k in range(20):
        g.Preconditioning(zs[k % pairs], r)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(600):
        g.Preconditioning(zs[k % pairs], r)
    torch.cuda.synchronize()
    print(f"{pairs} (r, z) pairs in rotation: {(time.perf_counter() - t0) / 600 * 1e6:.1f} us per apply (host wall clock)")
PY

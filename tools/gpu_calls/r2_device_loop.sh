#!/bin/bash
mkdir -p gpurun_out
timeout 60 ./build_tmp/cond; echo "conditional-node probe rc=$?"
timeout 900 python -m pytest tests/test_gpu_pcg.py -x -q > gpurun_out/r2_loop_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_loop_tests.log
for v in 0 1; do
  MAS_PCG_DEVICE_LOOP=$v timeout 300 python tools/profile_pcg.py 32 > gpurun_out/r2_loop${v}_pcg.txt 2> gpurun_out/r2_loop${v}_pcg.err
  echo "== device loop=$v rc=$?"; grep -E "median|iterations" gpurun_out/r2_loop${v}_pcg.txt
done
python - <<'PY'
import importlib, os, sys, time
import torch
sys.path.insert(0, os.getcwd())
pkg = importlib.import_module("preconditioner-for-cloth-and-deformable-body-simulation_b200")
S = pkg.synth
for n in (1024, 512, 64):
    mesh = S.cloth_rect_device(n, n, torch.device("cuda:0"))
    g = pkg.SeSchwarzPreconditioner(0)
    g.m_positions, g.m_neighbours = mesh.positions, (mesh.nbr_starts, mesh.nbr_idx)
    g.AllocatePrecoditioner(mesh.nv, 0, 0)
    g.PreparePreconditioner(mesh.diag, mesh.offdiag, mesh.nbr_starts)
    b = torch.from_numpy(S.residual(mesh.nv)).cuda()
    for loop in (0, 1, 0, 1):
        g.set_option(13, loop)
        pkg.pcg_solve(g, mesh.diag, mesh.offdiag, mesh.nbr_starts, mesh.nbr_idx, b)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            res = pkg.pcg_solve(g, mesh.diag, mesh.offdiag, mesh.nbr_starts, mesh.nbr_idx, b)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        print(f"cloth {n}x{n}: device loop {loop}: {res.iterations} iterations, {best * 1e3:.3f} ms, {best * 1e6 / res.iterations:.1f} us per iteration all-in, converged {res.converged}")
PY

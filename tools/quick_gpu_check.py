"""Scratch GPU sanity + timing run (development aid; bench.py is the contract)."""
import importlib, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("preconditioner-for-cloth-and-deformable-body-simulation_b200")
S = pkg.synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
t = time.time(); mesh = S.cloth(n); print("gen", time.time() - t, flush=True)
g = pkg.SeSchwarzPreconditioner(0)
t = time.time(); g.setup_from_mesh(mesh, device_inputs=True); torch.cuda.synchronize(); print("setup (first, incl. alloc)", time.time() - t)
print("levels", g.num_level, g.level_size().tolist(), "blocks", g.num_blocks, "prepare launches", g.prepare_launches)
d = g._dev_inputs
for i in range(3):
    torch.cuda.synchronize(); t = time.time()
    g.PreparePreconditioner(d[0], d[1], d[2])
    torch.cuda.synchronize(); print("prepare wall ms", (time.time() - t) * 1e3, "device ms", g.timing_ms(0))
r = torch.from_numpy(S.residual(mesh.nv)).cuda(); z = torch.empty_like(r)
for variant in [int(x) for x in os.environ.get('MAS_VARIANTS', '200').split(',')]:
    g.set_option(1, variant)
    for _ in range(5): g.Preconditioning(z, r)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 50
    e0.record()
    for _ in range(K): g.Preconditioning(z, r)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    B = g.num_blocks * 4656 * 4 + 32 * mesh.nv
    print(f"variant {variant}: apply {ms*1e3:.1f} us  -> {1e3/ms:.0f} applies/s, {B/ms/1e6:.0f} GB/s algorithmic, launches {g.apply_launches}")
print("z norm", float(z[:, :3].norm()))

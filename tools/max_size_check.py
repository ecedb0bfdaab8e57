"""Largest mesh the hierarchy supports: 33,554,432 vertices = 5 levels (the Int4 coarse table of the reference,
SeSchwarzPreconditioner.h:96, holds four ancestors; one vertex more needs a sixth level).  A 8192x4096 cloth is generated
on the device, set up and applied on one B200; prints one JSON line (sizes, times, the size-independent properties of
M^-1) and checks that 33,554,433 vertices are refused with MAS_ERR_UNSUPPORTED instead of overrunning the table.

    python tools/max_size_check.py [nx ny]
"""
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG_NAME = "preconditioner-for-cloth-and-deformable-body-simulation_b200"


def main():
    import torch
    pkg = importlib.import_module(PKG_NAME)
    S = pkg.synth
    nx, ny = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8192, 4096)
    dev = torch.device("cuda:0")
    out = {"nx": nx, "ny": ny}

    # one vertex over the limit: refused before anything is touched
    g = pkg.SeSchwarzPreconditioner(0)
    dummy = torch.zeros(8, dtype=torch.float32, device=dev)
    g.m_positions, g.m_neighbours = dummy, (dummy.to(torch.int32), dummy.to(torch.int32))
    try:
        g.AllocatePrecoditioner(32 ** 5 + 1, 0, 0)
        out["over_limit"] = "accepted (BUG)"
    except pkg.MasError as e:
        out["over_limit"] = str(e)
    g.close()

    t0 = time.perf_counter()
    mesh = S.cloth_rect_device(nx, ny, dev)
    torch.cuda.synchronize()
    out["generate_s"] = time.perf_counter() - t0
    nv = mesh.nv
    g = pkg.SeSchwarzPreconditioner(0)
    g.m_positions, g.m_neighbours = mesh.positions, (mesh.nbr_starts, mesh.nbr_idx)
    t0 = time.perf_counter()
    g.AllocatePrecoditioner(nv, 0, 0)
    torch.cuda.synchronize()
    out["allocate_ms"] = (time.perf_counter() - t0) * 1e3
    g.PreparePreconditioner(mesh.diag, mesh.offdiag, mesh.nbr_starts)
    g.PreparePreconditioner(mesh.diag, mesh.offdiag, mesh.nbr_starts)
    out["prepare_ms"] = g.timing_ms(0)
    out.update(nv=nv, num_level=g.num_level, level_size=g.level_size().tolist(), blocks=g.num_blocks,
               packed_inverse_gb=g.num_blocks * 18624 / 1e9)

    gen = torch.Generator(device=dev).manual_seed(1)
    r1 = torch.rand((nv, 4), generator=gen, device=dev) * 2 - 1
    r2 = torch.rand((nv, 4), generator=gen, device=dev) * 2 - 1
    r1[:, 3] = 0
    r2[:, 3] = 0
    z1, z2, z12, z1b = (torch.empty_like(r1) for _ in range(4))
    g.Preconditioning(z1, r1)
    g.Preconditioning(z2, r2)
    g.Preconditioning(z12, 2.0 * r1 - 0.5 * r2)
    g.Preconditioning(z1b, r1)
    torch.cuda.synchronize()
    d = lambda a, b: float((a[:, :3].double() * b[:, :3].double()).sum())
    lin = 2.0 * z1 - 0.5 * z2
    out["symmetry_rel"] = abs(d(r1, z2) - d(r2, z1)) / abs(d(r1, z1))
    out["positive"] = d(r1, z1) > 0 and d(r2, z2) > 0
    out["linearity_rel"] = float((z12 - lin)[:, :3].norm() / lin[:, :3].norm())
    out["w_zero"] = bool((z1[:, 3] == 0).all())
    out["deterministic"] = bool(torch.equal(z1, z1b))
    out["finite"] = bool(torch.isfinite(z1).all())

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        g.Preconditioning(z1, r1)
    e0.record()
    n = 20
    for _ in range(n):
        g.Preconditioning(z1, r1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    out["apply_ms"] = ms
    out["apply_algorithmic_gbs"] = (g.num_blocks * 18624 + 32 * nv) / (ms * 1e-3) / 1e9
    out["hbm_in_use_gb"] = (torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 1e9
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()

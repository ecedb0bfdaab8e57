"""Setup time of the two inversion kernels (MAS_OPT_INVERT_VARIANT 0 = tcgen05 tensor cores, the default; 1 = FP32 CUDA cores)
on the n x n cloth, and how far their z are apart.
    python tools/invert_variant_bench.py [n=1024] [variants, e.g. 1,0]
"""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    pkg = importlib.import_module("preconditioner-for-cloth-and-deformable-body-simulation_b200")
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    dev = torch.device("cuda:0")
    mesh = pkg.synth.cloth_rect_device(n, n, dev)
    out = {"nv": mesh.nv}
    ref = None
    variants = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 0]
    for variant in variants:
        g = pkg.SeSchwarzPreconditioner(0)
        g.set_option(8, variant)
        g.m_positions, g.m_neighbours = mesh.positions, (mesh.nbr_starts, mesh.nbr_idx)
        g.AllocatePrecoditioner(mesh.nv, 0, 0)
        ms = []
        for _ in range(6):
            g.PreparePreconditioner(mesh.diag, mesh.offdiag, mesh.nbr_starts)
            ms.append(g.timing_ms(0))
        out[f"variant{variant}_prepare_ms"] = min(ms[1:])
        r = torch.ones((mesh.nv, 4), device=dev)
        z = torch.empty_like(r)
        g.Preconditioning(z, r)
        torch.cuda.synchronize()
        if ref is None:
            ref = z.clone()
        else:
            out[f"variant{variant}_z_bit_identical"] = bool(torch.equal(ref, z))
            out[f"variant{variant}_z_rel_diff"] = float((z - ref)[:, :3].norm() / ref[:, :3].norm())
        g.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()

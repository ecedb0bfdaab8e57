import importlib, sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
pkg = importlib.import_module("preconditioner-for-cloth-and-deformable-body-simulation_b200")
from oracle import oracle_binding as ob
from helpers import bsr_matrix, cpu_pcg, make_oracle
S = pkg.synth
for name, mesh in (("tet16x16x8", S.tet_cube(16, 16, 8)), ("cloth50", S.cloth(50))):
    A = bsr_matrix(mesh); b = S.residual(mesh.nv)
    o = make_oracle(ob, mesh, "f"); o64 = make_oracle(ob, mesh, "d")
    g = pkg.SeSchwarzPreconditioner(0).setup_from_mesh(mesh)
    def gp(r):
        z = np.zeros_like(r); g.Preconditioning(z, r); return z
    print(name, "cpu loop + oracle32", cpu_pcg(A, b, o.apply)[1], "oracle64", cpu_pcg(A, b, o64.apply)[1], "gpu precond", cpu_pcg(A, b, gp)[1],
          "plain", cpu_pcg(A, b, None)[1])
    for mi in (1, 2, 3):
        res = pkg.pcg_solve(g, mesh.diag, mesh.offdiag, mesh.nbr_starts, mesh.nbr_idx, b, use_preconditioner=False, max_iter=mi)
        x_ref = cpu_pcg(A, b, None, max_iter=mi)[0]
        print("  plain it", mi, "x err", np.abs(res.x[:, :3] - x_ref).max() / np.abs(x_ref).max(), res.rel_residual)
    res = pkg.pcg_solve(g, mesh.diag, mesh.offdiag, mesh.nbr_starts, mesh.nbr_idx, b)
    resp = pkg.pcg_solve(g, mesh.diag, mesh.offdiag, mesh.nbr_starts, mesh.nbr_idx, b, use_preconditioner=False)
    print("  gpu harness: MAS", res.iterations, "plain", resp.iterations)
    for mi in (1, 2, 5, 20):
        res = pkg.pcg_solve(g, mesh.diag, mesh.offdiag, mesh.nbr_starts, mesh.nbr_idx, b, max_iter=mi)
        x_ref = cpu_pcg(A, b, gp, max_iter=mi)[0]
        print("  MAS it", mi, "x err", np.abs(res.x[:, :3] - x_ref).max() / np.abs(x_ref).max(), res.rel_residual)

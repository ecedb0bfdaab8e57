"""Development aid: a few PCG iterations on the 1M-vertex cloth (for ncu launch lists of the PCG kernels)."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("preconditioner-for-cloth-and-deformable-body-simulation_b200")
S = pkg.synth
mesh = S.cloth_rect_device(1024, 1024, torch.device("cuda:0"))
g = pkg.SeSchwarzPreconditioner(0)
g.m_positions, g.m_neighbours = mesh.positions, (mesh.nbr_starts, mesh.nbr_idx)
g.AllocatePrecoditioner(mesh.nv, 0, 0)
g.PreparePreconditioner(mesh.diag, mesh.offdiag, mesh.nbr_starts)
b = torch.from_numpy(S.residual(mesh.nv)).cuda()
it = int(sys.argv[1]) if len(sys.argv) > 1 else 4
res = pkg.pcg_solve(g, mesh.diag, mesh.offdiag, mesh.nbr_starts, mesh.nbr_idx, b, max_iter=it)
torch.cuda.synchronize()
print("iterations", res.iterations, "rel", res.rel_residual)

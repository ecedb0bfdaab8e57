import importlib, sys, os, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("preconditioner-for-cloth-and-deformable-body-simulation_b200")
S = pkg.synth
world = 2
mesh = S.cloth(96)
r = torch.from_numpy(S.residual(mesh.nv)).cuda()
streams = [torch.cuda.Stream() for _ in range(world)]
shards = [pkg.SeSchwarzPreconditioner(0, rank=k, world=world, stream=streams[k]) for k in range(world)]
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
d = dict(pos=dev(mesh.positions), st=dev(mesh.nbr_starts), ix=dev(mesh.nbr_idx), diag=dev(mesh.diag), off=dev(mesh.offdiag))
torch.cuda.synchronize()
for g in shards:
    g.m_positions, g.m_neighbours = d["pos"], (d["st"], d["ix"])
    g.AllocatePrecoditioner(mesh.nv, 0, 0)
arenas = [g.peer_local() for g in shards]
for g in shards:
    g.PreparePreconditioner(d["diag"], d["off"], d["st"], phase="begin")
torch.cuda.synchronize()
total = sum(g.exchange_tensor(0).clone() for g in shards)
for g in shards:
    g.exchange_tensor(0).copy_(total)
torch.cuda.synchronize()
for g in shards:
    g.prepare_end()
for g in shards:
    g.peer_attach(pointers=arenas)
for rep in range(3):
    zs = [torch.zeros_like(r) for _ in shards]
    torch.cuda.synchronize()
    for k, (g, z) in enumerate(zip(shards, zs)):
        t = time.time(); g.Preconditioning(z, r); print("rep", rep, "shard", k, "enqueue ms", (time.time() - t) * 1e3, flush=True)
    t = time.time(); torch.cuda.synchronize(); print("sync ms", (time.time() - t) * 1e3, "errors", [g.peer_error for g in shards], flush=True)

"""Development aid: per-kernel timeline of one PCG iteration on the 1M-vertex cloth (CUPTI through torch.profiler).
python tools/profile_pcg.py [iterations]"""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("preconditioner-for-cloth-and-deformable-body-simulation_b200")
S = pkg.synth
n = int(os.environ.get("MAS_N", 1024))
mesh = S.cloth_rect_device(n, n, torch.device("cuda:0"))
g = pkg.SeSchwarzPreconditioner(0)
g.m_positions, g.m_neighbours = mesh.positions, (mesh.nbr_starts, mesh.nbr_idx)
g.AllocatePrecoditioner(mesh.nv, 0, 0)
g.PreparePreconditioner(mesh.diag, mesh.offdiag, mesh.nbr_starts)
b = torch.from_numpy(S.residual(mesh.nv)).cuda()
it = int(sys.argv[1]) if len(sys.argv) > 1 else 32
if os.environ.get("MAS_PCG_PERSIST_L2") is not None:
    g.set_option(12, int(os.environ["MAS_PCG_PERSIST_L2"]))
solve = lambda: pkg.pcg_solve(g, mesh.diag, mesh.offdiag, mesh.nbr_starts, mesh.nbr_idx, b, max_iter=it)
solve(); torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    res = solve()
    torch.cuda.synchronize()
path = os.path.join(ROOT, "gpurun_out", "trace_pcg.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
short = lambda e: e["name"].replace("(anonymous namespace)::", "").split("(")[0].split("::")[-1]
first = [i for i, e in enumerate(ev) if short(e) == "spmv_dot_kernel"]
print("iterations", res.iterations, "rel", res.rel_residual, "spmv launches", len(first))
k = first[len(first) // 2]
k2 = first[len(first) // 2 + 1]
step = ev[k:k2]
base = step[0]["ts"]
for e in step:
    print(f"{short(e):28s} start {e['ts'] - base:8.1f} us  dur {e['dur']:7.1f} us  stream {e['args'].get('stream')}")
print("iteration period", ev[k2]["ts"] - base, "us; kernels per iteration", len(step))
per = [ev[first[i + 1]]["ts"] - ev[first[i]]["ts"] for i in range(len(first) - 1)]
per.sort()
print("median period", per[len(per) // 2], "min", per[0], "max", per[-1])
os.remove(path)

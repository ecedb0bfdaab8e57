"""Development aid: per-kernel timeline of one PreparePreconditioner (CUPTI through torch.profiler).
MAS_N=2048 python tools/profile_setup.py"""
import importlib, json, os, sys, collections
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("preconditioner-for-cloth-and-deformable-body-simulation_b200")
S = pkg.synth
n = int(os.environ.get("MAS_N", 1024))
g = pkg.SeSchwarzPreconditioner(0)
if os.environ.get("MAS_CONFIG") is not None:          # a BASELINE config (synth.config), inputs resident on the device
    mesh = S.config(int(os.environ["MAS_CONFIG"]), proximity=bool(os.environ.get("MAS_PROXIMITY")))
    g.setup_from_mesh(mesh, device_inputs=True)
    d = g._dev_inputs
    prepare = lambda: g.PreparePreconditioner(d[0], d[1], d[2], d[3], d[4], d[5], mesh.ef_total, mesh.ee_total, mesh.vf_total)
    label = f"config {os.environ['MAS_CONFIG']} ({mesh.nv} verts)"
else:
    mesh = S.cloth_rect_device(n, n, torch.device("cuda:0"))
    g.m_positions, g.m_neighbours = mesh.positions, (mesh.nbr_starts, mesh.nbr_idx)
    g.AllocatePrecoditioner(mesh.nv, 0, 0)
    prepare = lambda: g.PreparePreconditioner(mesh.diag, mesh.offdiag, mesh.nbr_starts)
    label = f"cloth {n}x{n}"
for _ in range(3):
    prepare()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    prepare()
    torch.cuda.synchronize()
path = os.path.join(ROOT, "gpurun_out", "trace_setup.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
ev.sort(key=lambda e: e["ts"])
short = lambda e: e["name"].replace("(anonymous namespace)::", "").split("(")[0].split("::")[-1][:40]
t0 = ev[0]["ts"]
print(f"{label}: prepare device time {g.timing_ms(0):.3f} ms; {len(ev)} GPU activities")
for e in ev:
    if e["dur"] >= 15:
        print(f"{short(e):40s} start {e['ts'] - t0:9.1f} us  dur {e['dur']:8.1f} us")
tot = collections.Counter()
for e in ev:
    tot[short(e)] += e["dur"]
print("sum by kernel:", ", ".join(f"{k} {v:.0f}" for k, v in tot.most_common(12)))
print("span", ev[-1]["ts"] + ev[-1]["dur"] - t0, "us")
os.remove(path)

"""Development aid: per-kernel timeline of the sharded apply on rank 0 via torch.profiler (CUPTI sees our kernels too).
torchrun --nproc-per-node N tools/profile_sharded.py [variant]"""
import importlib, json, os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("preconditioner-for-cloth-and-deformable-body-simulation_b200")
S = pkg.synth
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
mesh = S.cloth_rect_device(int(os.environ.get("MAS_NX", os.environ.get("MAS_N", 1024))), int(os.environ.get("MAS_NY", os.environ.get("MAS_N", 1024))), torch.device(f"cuda:{local}"))
g = pkg.SeSchwarzPreconditioner(device=local, rank=rank, world=world, stream=torch.cuda.current_stream())
if len(sys.argv) > 1:
    g.set_option(1, int(sys.argv[1]))
t = lambda a: a.cuda() if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a)).cuda()
g.m_positions, g.m_neighbours = t(mesh.positions), (t(mesh.nbr_starts), t(mesh.nbr_idx))
g.AllocatePrecoditioner(mesh.nv, 0, 0)
d = (t(mesh.diag), t(mesh.offdiag), t(mesh.nbr_starts))
if world > 1:
    drv = pkg.partition.ShardedSchwarzPreconditioner(g)
    assert drv.attach_peers()
    drv.PreparePreconditioner(*d)
else:
    g.PreparePreconditioner(*d)
r = t(S.residual(mesh.nv)); z = torch.zeros_like(r)
for _ in range(20):
    g.Preconditioning(z, r)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(10):
        g.Preconditioning(z, r)
    torch.cuda.synchronize()
if world > 1:
    dist.barrier()
if rank in (0, world - 1):
    path = os.path.join(ROOT, "gpurun_out", f"trace_w{world}_r{rank}.json")
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
    ev.sort(key=lambda e: e["ts"])
    # split into steps at restrict_fine / first kernel of each graph launch
    names = [e["name"].split("(")[0].split("::")[-1] for e in ev]
    per = len(ev) // 10
    print(f"---- rank {rank}")
    step = ev[per * 5: per * 6]
    t0 = min(e["ts"] for e in step)
    for e in step:
        print(f"{e['name'].replace('(anonymous namespace)::','').split('(')[0].split('::')[-1]:28s} start {e['ts'] - t0:8.1f} us  dur {e['dur']:7.1f} us  stream {e['args'].get('stream')}")
    print("step span", max(e["ts"] + e["dur"] for e in step) - t0, "us; kernels per step", per)
    nxt = ev[per * 6]["ts"] - t0 if len(ev) > per * 6 else None
    print("next step starts at", nxt)
if world > 1:
    dist.destroy_process_group()

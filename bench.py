#!/usr/bin/env python
"""bench.py — Preconditioning() applies/s (+ setup ms) of the MAS preconditioner on a 1M-vertex cloth.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config I]

One "step" = one Preconditioning() apply (the per-PCG-iteration hot path, SeSchwarzPreconditioner.cpp:100-110) over
the whole synthetic mesh.  N>1 is launched by torchrun (one rank per GPU); the 32-node fine domains are sharded in
Morton-contiguous ranges and the only data-path exchange per apply is an all-gather of the level-2 residuals over peer
memory inside the apply graph (NCCL all-reduce with --nccl-exchange), plus one NCCL all-reduce of the coarse Galerkin
accumulators per setup.  Scaling (N>1, default config): "weak" (default) shards ONE
cloth of N x 1,048,576 vertices (1024x1024, 2048x1024, 2048x2048, 4096x2048 for N = 1, 2, 4, 8), i.e. per-GPU work is
fixed at the 1M-vertex cloth the metric is quoted on and `value` counts 1M-vertex applies/s over all ranks
(N x whole-mesh applies/s); "--scaling strong" keeps the 1M-vertex mesh and splits it N ways (latency-bound from N = 4).

Keys beyond the base contract:
  value          applies/s with r and z resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e            same metric through the public host-pointer API: pinned host r -> H2D -> apply -> D2H z, every step
                 (timed with both host stagings the library offers, copy engine and MAS_OPT_HOST_PULL; the better is reported)
  setup_ms       PreparePreconditioner() per call, device-resident inputs (CUDA events inside the library)
  roofline       dominant kernel (level-0 solve): algorithmic bytes / its CUDA-event duration vs measured HBM peak
  cpu_baseline   the reference's own CPU code (oracle/_ref, all host threads) on the same mesh; rank 0, N=1 only
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG_NAME = "preconditioner-for-cloth-and-deformable-body-simulation_b200"
METRIC = "Preconditioning() applies/s, 1M-vert cloth"
WORKLOADS = {0: "cloth 64x64 (4,096 verts)", 1: "cloth 512x512 (262,144 verts) + EF/EE/VF stencils",
             2: "cloth 1024x1024 (1,048,576 verts), 8-neighbour springs, 4 levels",
             3: "tet cube 128x128x64 (1,048,576 verts)", 4: "cloth 2048x2048 (4,194,304 verts), 5 levels"}


def weak_grid(world: int):
    """(a, b): the weak-scaling cloth is 1024a x 1024b vertices with a * b = world."""
    b = 1
    while (b * 2) * (b * 2) <= world and world % (b * 2) == 0:
        b *= 2
    return world // b, b


def workload_and_config(args, world: int):
    """`config` of the JSON line: identical for both arms (the driver compares them), so only what is known before anything
    runs.  Returns (config dict, units = 1M-vertex applies per whole-mesh apply, nv)."""
    weak = world > 1 and args.scaling == "weak" and args.config == 2
    workload = WORKLOADS[args.config]
    nv = {0: 4096, 1: 262144, 2: 1048576, 3: 1048576, 4: 4194304}[args.config]
    units = 1.0
    if weak:
        a, b = weak_grid(world)
        nv = 1048576 * world
        units = float(world)
        workload = (f"cloth {1024 * a}x{1024 * b} ({nv:,} verts = {world} x 1,048,576), 8-neighbour springs, Morton-sharded over "
                    f"{world} GPUs; the reference arm (one host, no GPUs) times ONE 1,048,576-vertex unit on all host threads")
    elif world > 1:
        workload += f", Morton-sharded over {world} GPUs (strong scaling); the reference arm times the same mesh on the host"
    cfg = {"workload": workload, "nv": nv,
           "units": (f"value = {units:g} x whole-mesh applies/s: one apply of the sharded {nv:,}-vertex mesh counts as {units:g} applies of "
                     f"the 1M-vertex cloth (per-GPU work fixed)") if weak else "whole-mesh applies/s",
           "l2": "inputs larger than L2: 630 MB of packed inverses streamed per step per GPU vs 126 MB L2 (the CPU arm streams the "
                 "same 630 MB per 1M vertices from host memory)",
           "timing": "ours: CUDA events on the launching stream, max over ranks; reference: host wall clock, mean over steps",
           "parallelism": f"morton-sharded x{world}" if world > 1 else "single GPU"}
    return cfg, units, nv


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def algorithmic_bytes(n_blocks: int, nv: int) -> int:
    """packed symmetric FP32 inverses read once + 16 B residual read + 16 B z write per vertex."""
    return n_blocks * 4656 * 4 + 32 * nv


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_peak_tflops():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["bf16_tflops"])
    except Exception:
        return 1680.0          # this pool's measured dense bf16 figure (B200_PROFILING.md fallback)


def traffic_from_profile():
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture, if present."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)).get("solve_fine_dram_bytes_per_launch")
        except Exception:
            return None
    return None


# ----------------------------------------------------------------------------------------------------------------------
def cpu_reference_apply(mesh, r, steps: int, warmup: int, threads: int):
    """Times the reference's own CPU implementation (oracle/_ref = SeSchwarzPreconditioner.cpp compiled here) or, if
    that shared object did not travel, the plain-C oracle port.  Returns (applies/s, setup_ms, kind)."""
    from oracle import ref_binding as rb
    # beyond 33,792 level-1 nodes (nv > 1,081,344) the stock reference truncates a scan (Q5, cpp:989-994) and then overruns
    # its own buffers (observed: segfault at 2048^2): such meshes are timed on the build with that four-line cull removed
    big = mesh.nv > 33792 * 32
    if rb.available(q5fix=big) and not os.environ.get("MAS_BENCH_FORCE_PORT"):
        p = rb.RefPreconditioner(threads=threads, q5fix=big)
        kind = "reference"
    else:
        from oracle import oracle_binding as ob
        os.environ.setdefault("OMP_NUM_THREADS", str(threads))
        p = ob.OraclePreconditioner("f")
        kind = "port"
    p.allocate(mesh)
    p.prepare()                       # first call page-faults the buffers in
    t0 = time.perf_counter()
    p.prepare()
    setup_ms = (time.perf_counter() - t0) * 1e3
    import numpy as np
    z = np.zeros_like(r)
    for _ in range(max(1, warmup)):
        p.apply(r, out=z)
    best = float("inf")
    t_all = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        p.apply(r, out=z)
        best = min(best, time.perf_counter() - t0)
    mean = (time.perf_counter() - t_all) / steps
    cpu_reference_apply.last_z = z          # the reference's z on this (mesh, r): value parity is checked against it
    return 1.0 / mean, 1.0 / best, setup_ms, kind


def _mesh_to_npz(mesh, r, path):
    import numpy as np
    np.savez(path, name=mesh.name, nv=mesh.nv, positions=mesh.positions, nbr_starts=mesh.nbr_starts, nbr_idx=mesh.nbr_idx, diag=mesh.diag,
             offdiag=mesh.offdiag, edges=mesh.edges, faces=mesh.faces, ef=mesh.ef, ee=mesh.ee, vf=mesh.vf,
             totals=np.array([mesh.ef_total, mesh.ee_total, mesh.vf_total]), r=r)


def _mesh_from_npz(path):
    import numpy as np
    pkg = importlib.import_module(PKG_NAME)
    d = np.load(path)
    m = pkg.synth.Mesh(str(d["name"]), int(d["nv"]), d["positions"], d["nbr_starts"], d["nbr_idx"], d["diag"], d["offdiag"], d["edges"], d["faces"],
                       d["ef"], d["ee"], d["vf"], int(d["totals"][0]), int(d["totals"][1]), int(d["totals"][2]))
    return m, d["r"]


def cpu_leg_child(path, steps, warmup, threads, force_port):
    """Internal (`bench.py --cpu-leg-child`): the CPU reference leg in a process of its own, so that a crash of the reference's
    code (it overruns its fixed allocations on some inputs) cannot take the bench line down."""
    import numpy as np
    mesh, r = _mesh_from_npz(path)
    if force_port:
        os.environ["MAS_BENCH_FORCE_PORT"] = "1"
    mean_rate, best_rate, setup_ms, kind = cpu_reference_apply(mesh, r, steps, warmup, threads)
    np.save(path + ".z.npy", cpu_reference_apply.last_z)
    print(json.dumps({"mean_rate": mean_rate, "best_rate": best_rate, "setup_ms": setup_ms, "kind": kind}), flush=True)


def cpu_leg_isolated(mesh, r, steps, warmup, threads):
    """Runs the CPU reference leg in a child process.  Returns (mean_rate, best_rate, setup_ms, kind, z_ref, note)."""
    import tempfile
    import numpy as np
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "mesh.npz")
        _mesh_to_npz(mesh, r, path)
        note = None
        for force_port in (False, True):
            cmd = [sys.executable, os.path.abspath(__file__), "--cpu-leg-child", path, "--steps", str(steps), "--warmup", str(warmup)]
            if force_port:
                cmd.append("--cpu-leg-port")
            p = subprocess.run(cmd, capture_output=True, text=True, timeout=1800, env=dict(os.environ, MAS_BENCH_THREADS=str(threads)))
            if p.returncode == 0 and p.stdout.strip():
                res = json.loads(p.stdout.strip().splitlines()[-1])
                z = np.load(path + ".z.npy")
                return res["mean_rate"], res["best_rate"], res["setup_ms"], res["kind"], z, note
            note = (f"the reference's own code crashed on this input (exit code {p.returncode}: it overruns its fixed allocations, SURVEY Q6); "
                    "timed the plain-C port of the same algorithm instead")
        raise RuntimeError("CPU reference leg failed twice: " + (p.stderr or "")[-300:])


def cpu_reference_pcg(mesh, b, threads: int):
    """PCG to 1e-5 on the host: scipy block-CSR SpMV + the reference's own Preconditioning() (oracle/_ref), all threads."""
    from oracle import ref_binding as rb
    from oracle.cpu_pcg import bsr_matrix, cpu_pcg
    big = mesh.nv > 33792 * 32
    if rb.available(q5fix=big):
        p = rb.RefPreconditioner(threads=threads, q5fix=big)
        kind = "reference"
    else:
        from oracle import oracle_binding as ob
        p = ob.OraclePreconditioner("f")
        kind = "port"
    p.allocate(mesh)
    p.prepare()
    A = bsr_matrix(mesh)
    t0 = time.perf_counter()
    _, its = cpu_pcg(A, b, p.apply)
    return {"iterations": its, "solve_ms": (time.perf_counter() - t0) * 1e3, "kind": kind, "cores": threads, "tol": 1e-5}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    pkg = importlib.import_module(PKG_NAME)
    cfg, units, _ = workload_and_config(args, max(world, args.gpus))
    mesh = pkg.synth.config(args.config, proximity=getattr(args, "proximity", False))      # one unit on all host threads
    r = pkg.synth.residual(mesh.nv)
    threads = host_threads()
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    t0 = time.perf_counter()
    mean_rate, best_rate, setup_ms, kind, _, note = cpu_leg_isolated(mesh, r, steps, warmup, threads)
    if kind != "reference":
        threads = 1
    out = {
        "impl": "reference", "metric": METRIC, "value": mean_rate, "unit": "applies/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": 1e3 / mean_rate, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "setup_ms": setup_ms,
        "cpu_baseline": {"value": mean_rate, "unit": "applies/s", "cores": threads, "kind": kind,
                         "sample": f"whole {WORKLOADS[args.config]} mesh, {steps} applies after 1 setup + {warmup} warm-up applies; "
                                   f"best single apply {1e3 / best_rate:.2f} ms"},
        "e2e": {"value": mean_rate, "unit": "applies/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    if note:
        out["cpu_baseline"]["note"] = note
    print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout for the one JSON line
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    pkg = importlib.import_module(PKG_NAME)
    S = pkg.synth
    dev = f"cuda:{local}"
    weak = world > 1 and args.scaling == "weak" and args.config == 2
    cfg, units, _ = workload_and_config(args, world)
    if args.config == 2:
        # generated on the device (bit-identical to the numpy recipe, tests/test_synth.py): no multi-GB host pass per rank
        a, b = weak_grid(world) if weak else (1, 1)
        mesh = S.cloth_rect_device(1024 * a, 1024 * b, torch.device(dev))
    elif args.config == 4:
        mesh = S.cloth_rect_device(2048, 2048, torch.device(dev))      # strong scaling: the 4.2M-vertex cloth split N ways
    elif args.config == 1 and args.proximity:
        mesh = S.config(1, proximity=True)
        cfg["workload"] = f"cloth 512x512 (262,144 verts) folded in half + proximity EF/EE/VF stencils ({mesh.ef_total}/{mesh.ee_total}/{mesh.vf_total})"
    else:
        mesh = S.config(args.config)
    nv = mesh.nv
    assert nv == cfg["nv"]

    stream = torch.cuda.current_stream()
    g = pkg.SeSchwarzPreconditioner(device=local, rank=rank, world=world, stream=stream)
    t = lambda a: a.to(dev) if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    tb = lambda a: torch.from_numpy(np.frombuffer(np.ascontiguousarray(a).tobytes(), np.uint8).copy()).to(dev)
    g.m_positions = t(mesh.positions)
    g.m_edges = t(mesh.edges) if mesh.ne else None
    g.m_faces = t(mesh.faces) if mesh.nf else None
    g.m_neighbours = (t(mesh.nbr_starts), t(mesh.nbr_idx))
    g.AllocatePrecoditioner(nv, mesh.ne, mesh.nf)
    d_in = (t(mesh.diag), t(mesh.offdiag), t(mesh.nbr_starts), tb(mesh.ef) if mesh.ef.size else None,
            tb(mesh.ee) if mesh.ee.size else None, tb(mesh.vf) if mesh.vf.size else None)

    def prepare(eng=None):
        eng = eng or g
        if eng.world == 1:
            eng.PreparePreconditioner(d_in[0], d_in[1], d_in[2], d_in[3], d_in[4], d_in[5], mesh.ef_total, mesh.ee_total, mesh.vf_total)
        else:
            eng.PreparePreconditioner(d_in[0], d_in[1], d_in[2], d_in[3], d_in[4], d_in[5], mesh.ef_total, mesh.ee_total,
                                      mesh.vf_total, phase="begin")
            dist.all_reduce(eng.exchange_tensor(0))
            eng.prepare_end()

    if args.variant is not None:
        g.set_option(1, args.variant)
    if args.invert_variant:
        g.set_option(8, args.invert_variant)       # MAS_OPT_INVERT_VARIANT: 0 tensor cores (default), 1 FP32 CUDA cores
    p2p = False
    if world > 1:
        drv = pkg.partition.ShardedSchwarzPreconditioner(g)
        p2p = drv.attach_peers() and not args.nccl_exchange
        if not p2p:
            drv.p2p = False
    prepare()
    setup_wall = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        prepare()
        torch.cuda.synchronize()
        setup_wall.append((time.perf_counter() - t0) * 1e3)
    setup_ms = min(setup_wall)
    setup_device_ms = g.timing_ms(0) if world == 1 else None
    # the same setup with the clustering rebuilt every time (MAS_OPT_CACHE_HIERARCHY = 0): what a first prepare, or one with
    # new collision stencils, costs
    setup_full_ms = None
    if world == 1 and not args.lean:
        g.set_option(10, 0)
        acc = []
        for _ in range(3):
            prepare()
            acc.append(g.timing_ms(0))
        setup_full_ms = min(acc)
        g.set_option(10, 1)
        prepare()

    r = t(S.residual(nv))
    z = torch.zeros_like(r)
    exch = g.exchange_tensor(1) if world > 1 else None

    def step():
        if world == 1 or p2p:
            g.Preconditioning(z, r)          # one CUDA graph per rank; sharded ranks exchange over peer memory on the device
        else:
            g.apply_begin(r)
            dist.all_reduce(exch)
            g.apply_end(z)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    warmup = max(3, args.warmup)
    for _ in range(warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    tms = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_per_step = float(tms.item()) / args.steps
    if not args.lean:
        # hold the load ~0.4 s longer so that the 50 ms sampler sees the clocks under load (the timed region of 200 applies
        # lasts 21 ms).  Every rank runs the same number of extra steps: ms_per_step is the all-reduced maximum.
        for _ in range(int(min(20000, max(50, 400.0 / max(ms_per_step, 1e-3))))):
            step()
        barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches_per_step = g.apply_launches

    # ---- sharded runs: the timed z is checked in the run.  (1) no device-side peer wait timed out on any rank (that would
    # leave stale coarse residuals behind: mas_synchronize raises); (2) the shards merge to the z a single GPU computes for the
    # same mesh and r ("G-GPU z equals 1-GPU z"; only the summation order of the FP64 coarse accumulators differs).
    parity = None
    if world > 1:
        g.synchronize()
        peer_err = torch.tensor([g.peer_error], device=dev, dtype=torch.int32)
        dist.all_reduce(peer_err, op=dist.ReduceOp.MAX)
        b0, b1 = g.owned_fine_blocks
        s2o = torch.from_numpy(g.sorted_get_original().astype(np.int64)).to(dev)
        own = s2o[min(32 * b0, nv):min(32 * b1, nv)]
        merged = torch.zeros_like(z)
        merged[own] = z[own]
        covered = torch.zeros(nv, device=dev, dtype=torch.int32)
        covered[own] = 1
        dist.all_reduce(merged)
        dist.all_reduce(covered)
        parity = {"peer_error": int(peer_err.item()), "every_vertex_owned_once": bool((covered == 1).all().item())}
        if rank == 0:
            one = pkg.SeSchwarzPreconditioner(device=local, stream=stream)
            one.m_positions, one.m_edges, one.m_faces, one.m_neighbours = g.m_positions, g.m_edges, g.m_faces, g.m_neighbours
            if args.invert_variant:
                one.set_option(8, args.invert_variant)
            one.AllocatePrecoditioner(nv, mesh.ne, mesh.nf)
            prepare(one)
            z1 = torch.empty_like(r)
            one.Preconditioning(z1, r)
            torch.cuda.synchronize()
            rel = float((merged - z1)[:, :3].double().norm() / z1[:, :3].double().norm())
            parity.update({"rel_l2_sharded_vs_single_gpu": rel, "tolerance": 1e-5, "ok": bool(rel < 1e-5 and parity["peer_error"] == 0
                                                                                               and parity["every_vertex_owned_once"])})
            one.close()
            del z1
        del merged, covered

    # ---- BASELINE config 4 beside the headline (its strong-scaling curve rides on the driver's 1/2/4/8-GPU runs of the default
    # command): the 2048^2 cloth (4,194,304 vertices, five levels) split over the same N ranks, HBM-resident applies/s
    strong4 = None
    if args.config == 2 and not args.lean and not args.no_strong:
        try:
            m4 = S.cloth_rect_device(2048, 2048, torch.device(dev))
            e4 = pkg.SeSchwarzPreconditioner(device=local, rank=rank, world=world, stream=stream)
            e4.m_positions, e4.m_neighbours = m4.positions, (m4.nbr_starts, m4.nbr_idx)
            e4.AllocatePrecoditioner(m4.nv, 0, 0)
            p2p4 = False
            if world > 1:
                p2p4 = pkg.partition.ShardedSchwarzPreconditioner(e4).attach_peers()
                e4.PreparePreconditioner(m4.diag, m4.offdiag, m4.nbr_starts, phase="begin")
                dist.all_reduce(e4.exchange_tensor(0))
                e4.prepare_end()
            else:
                e4.PreparePreconditioner(m4.diag, m4.offdiag, m4.nbr_starts)
            if world == 1 or p2p4:
                r4 = t(S.residual(m4.nv))
                z4 = torch.zeros_like(r4)
                for _ in range(5):
                    e4.Preconditioning(z4, r4)
                barrier()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n4 = max(20, min(args.steps, 100))
                a0.record()
                for _ in range(n4):
                    e4.Preconditioning(z4, r4)
                a1.record()
                barrier()
                t4 = torch.tensor([a0.elapsed_time(a1) / n4], device=dev, dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(t4, op=dist.ReduceOp.MAX)
                    e4.synchronize()
                strong4 = {"workload": WORKLOADS[4] + (f", Morton-sharded over {world} GPUs" if world > 1 else ""), "n_gpus": world,
                           "applies_per_s": 1e3 / float(t4.item()), "ms_per_apply": float(t4.item()), "steps": n4,
                           "scaling": "strong", "levels": e4.level_size().tolist()}
                del r4, z4
            e4.close()
            del m4, e4
            torch.cuda.empty_cache()
        except Exception as exc:                      # noqa: BLE001
            strong4 = {"error": repr(exc)}

    # ---- e2e: public host-pointer API, pinned host buffers, H2D + apply + D2H inside the timed region
    r_np = S.residual(nv)
    r_h = torch.from_numpy(r_np).pin_memory()
    z_h = torch.zeros_like(r_h).pin_memory()
    r_stage, z_stage = torch.empty_like(r), torch.empty_like(r)

    def e2e_step(zb=None, rb=None):
        zb, rb = (z_h if zb is None else zb), (r_h if rb is None else rb)
        if world == 1 or p2p:
            g.Preconditioning(zb, rb)            # mas_apply(MAS_MEM_HOST): host r in, apply graph, host z out, synchronous
        else:
            r_stage.copy_(rb, non_blocking=True)
            g.apply_begin(r_stage)
            dist.all_reduce(exch)
            g.apply_end(z_stage)
            zb.copy_(z_stage, non_blocking=True)
            torch.cuda.current_stream().synchronize()

    def time_e2e(n, zb=None, rb=None, warm=3):
        for _ in range(warm):
            e2e_step(zb, rb)
        barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            e2e_step(zb, rb)
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / n], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return 1.0 / float(dt.item())

    e2e_steps = 1 if args.lean else max(5, min(args.steps, 50))
    e2e_rate = time_e2e(e2e_steps, warm=0 if args.lean else 3)
    if world == 1 or p2p:
        h2d, d2h = g.get_int(16), g.get_int(17)            # counted by the library from the copies it issued
    else:
        h2d = d2h = 16 * nv
    if world > 1:
        moved = torch.tensor([h2d, d2h], device=dev, dtype=torch.int64)
        dist.all_reduce(moved)
        h2d, d2h = int(moved[0].item()), int(moved[1].item())

    # ---- dominant kernel alone, CUDA events on the launching stream (inside the library), un-captured launches
    fine_ms = None
    if world == 1 and not args.lean:
        g.set_option(3, 1)
        for _ in range(3):
            g.Preconditioning(z, r)
        acc = []
        for _ in range(max(10, min(args.steps, 100))):
            g.Preconditioning(z, r)
            acc.append(g.timing_ms(2))
        g.set_option(3, 0)
        acc.sort()
        fine_ms = sum(acc) / len(acc)
        fine_ms_med = acc[len(acc) // 2]

    # ---- sharded runs: every rank's own level-0 kernel time (same banks per rank), to tell GPU-to-GPU spread from exchange cost
    per_rank_fine_ms = None
    if world > 1 and p2p and not args.lean:
        g.set_option(3, 1)
        for _ in range(3):
            g.Preconditioning(z, r)
        acc = []
        for _ in range(30):
            g.Preconditioning(z, r)
            acc.append(g.timing_ms(2))
        g.set_option(3, 0)
        mine = torch.tensor([sum(acc) / len(acc)], device=dev, dtype=torch.float64)
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        per_rank_fine_ms = [round(float(x.item()), 5) for x in gathered]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    if torch.is_tensor(mesh.positions) and world == 1:
        # host copy for the CPU reference leg
        for name in ("positions", "nbr_starts", "nbr_idx", "diag", "offdiag"):
            setattr(mesh, name, getattr(mesh, name).cpu().numpy())
    n_blocks = g.num_blocks
    lv = g.level_size().tolist()
    n_fine = ((nv + 31) // 32)
    peak, peak_src = measured_peak_gbs()
    roof = None
    if fine_ms:
        fine_bytes = algorithmic_bytes(n_fine, nv)
        achieved = fine_bytes / (fine_ms * 1e-3) / 1e9
        # setup: bytes that have to move (Hessian blocks and indices in, packed inverses out) and the tensor-core work of the
        # inversion (six rank-16 updates of 128 x 96 per system, three TF32 products each)
        nnz = int(mesh.nbr_starts[-1])
        setup_bytes = n_blocks * 4656 * 4 + nv * (36 + 16 + 8) + nnz * (36 + 4)
        setup_flop = n_blocks * 6 * 3 * 2 * 128 * 96 * 16
        tf32_peak = measured_peak_tflops() / 2.0
        roof = {"bound": "hbm", "kernel": "solve_fine_kernel (level-0 SchwarzLocalXSym fused with gather/prolongation)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": fine_bytes, "kernel_ms_mean": fine_ms, "kernel_ms_median": fine_ms_med,
                "traffic": traffic_from_profile(),
                "traffic_source": "profiles/roofline_traffic.json (ncu --set full of this kernel, dram__bytes_read + dram__bytes_write; not re-measured in this run)",
                "whole_apply": {"algorithmic_bytes": algorithmic_bytes(n_blocks, nv),
                                "achieved": algorithmic_bytes(n_blocks, nv) / (ms_per_step * 1e-3) / 1e9,
                                "frac": algorithmic_bytes(n_blocks, nv) / (ms_per_step * 1e-3) / 1e9 / peak,
                                "frac_of_8TBs": algorithmic_bytes(n_blocks, nv) / (ms_per_step * 1e-3) / 8e12},
                "setup": {"ms": setup_device_ms, "bytes_bound": setup_bytes, "hbm_gbs": setup_bytes / (setup_device_ms * 1e-3) / 1e9,
                          "hbm_frac": setup_bytes / (setup_device_ms * 1e-3) / 1e9 / peak,
                          "tensor_flop": setup_flop, "tensor_tflops": setup_flop / (setup_device_ms * 1e-3) / 1e12,
                          "tensor_peak_tf32_tflops": tf32_peak, "tensor_frac": setup_flop / (setup_device_ms * 1e-3) / 1e12 / tf32_peak,
                          "note": "whole PreparePreconditioner (device time); tf32 peak = half the measured dense bf16 peak; the kernel is "
                                  "bound by its serial per-panel chain (pivot inverse -> operand store -> MMA), not by either roof"}}

    # ---- PCG to 1e-5 with MAS through the device harness (BASELINE config 2: iteration count and wall time)
    pcg = None
    if world == 1 and not args.lean:
        idx_d = t(mesh.nbr_idx)
        b_d = t(S.residual(nv))
        pkg.pcg_solve(g, d_in[0], d_in[1], d_in[2], idx_d, b_d)          # warm-up (graph instantiate, workspace)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = pkg.pcg_solve(g, d_in[0], d_in[1], d_in[2], idx_d, b_d)
        torch.cuda.synchronize()
        pcg_ms = (time.perf_counter() - t0) * 1e3
        plain = pkg.pcg_solve(g, d_in[0], d_in[1], d_in[2], idx_d, b_d, use_preconditioner=False, max_iter=5000)
        # the iteration alone: the difference between two solves capped at k and 2k iterations (the per-solve work - the ELL
        # image of A, graph capture, the first apply - cancels)
        steady_us = None
        k = max(4, min(32, res.iterations // 2))
        if res.iterations >= 2 * k:
            def capped(n):
                best = 1e30
                for _ in range(3):
                    t0 = time.perf_counter()
                    pkg.pcg_solve(g, d_in[0], d_in[1], d_in[2], idx_d, b_d, max_iter=n)
                    torch.cuda.synchronize()
                    best = min(best, time.perf_counter() - t0)
                return best
            steady_us = (capped(2 * k) - capped(k)) / k * 1e6
        pcg = {"tol": 1e-5, "iterations": res.iterations, "converged": res.converged, "rel_residual": res.rel_residual,
               "solve_ms": pcg_ms, "ms_per_iteration": pcg_ms / max(1, res.iterations),
               "us_per_iteration_steady": steady_us,
               "launches_per_iteration": res.launches_per_iteration, "iterations_unpreconditioned": plain.iterations,
               "timing": "host wall clock around mas_pcg_solve, device-resident A/b/x; ms_per_iteration = whole solve / iterations "
                         "(includes building the ELL image of A, graph capture and the first apply); us_per_iteration_steady = "
                         "(solve capped at 2k iterations - solve capped at k) / k, best of 3 each"}

    cpu = None
    if world == 1 and not args.no_cpu_baseline and not args.lean:
        threads = host_threads()
        n_cpu = 10
        mean_rate, best_rate, cpu_setup_ms, kind, z_ref, cpu_note = cpu_leg_isolated(mesh, r_np, n_cpu, 1, threads)
        cpu = {"value": mean_rate, "unit": "applies/s", "cores": threads if kind == "reference" else 1, "kind": kind,
               "sample": f"whole {cfg['workload']} mesh: 1 setup + 1 warm-up + {n_cpu} timed applies, " +
                         ("all host threads" if kind == "reference" else "one thread (scalar port)") + ", in a child process",
               "best_ms": 1e3 / best_rate, "setup_ms": cpu_setup_ms}
        if cpu_note:
            cpu["note"] = cpu_note
        # value parity at full size: the z of the timed GPU path against the reference's z on the same mesh and r, and both
        # against FP64 arithmetic (the plain-C restatement built in double: the arbiter of SURVEY 8c)
        g.Preconditioning(z_h, r_h)
        z_gpu = z_h.numpy()
        rel = lambda a, b: float(np.linalg.norm((a - b)[:, :3].astype(np.float64)) / np.linalg.norm(b[:, :3].astype(np.float64)))
        parity = {"rel_l2_gpu_vs_reference": rel(z_gpu, z_ref), "reference_kind": kind}
        if not args.no_arbiter:
            try:
                from oracle import oracle_binding as ob
                t0 = time.perf_counter()
                o64 = ob.OraclePreconditioner("d")
                o64.allocate(mesh)
                o64.prepare()
                z64 = o64.apply(r_np)
                parity.update({"rel_l2_gpu_vs_f64": rel(z_gpu, z64), "rel_l2_reference_vs_f64": rel(z_ref, z64),
                               "arbiter": "plain-C restatement in double precision (oracle/mas_oracle.c)",
                               "arbiter_s": time.perf_counter() - t0})
                parity["ok"] = bool(parity["rel_l2_gpu_vs_f64"] <= 8 * parity["rel_l2_reference_vs_f64"] + 1e-6)
                parity["tolerance"] = "|z_gpu - z_f64| <= 8 |z_ref - z_f64| + 1e-6 |z_f64| (tensor-core setup; 2x with --invert-variant 1)"
                del o64
            except Exception as exc:                      # noqa: BLE001
                parity["arbiter_error"] = repr(exc)
        if args.cpu_pcg:
            cpu["pcg"] = cpu_reference_pcg(mesh, r_np, threads)

    if world > 1:
        # per-rank share of the algorithmic bytes against one GPU's HBM peak (whole apply; no kernel-only timing here)
        b0, b1 = g.owned_fine_blocks
        own_verts = max(0, min(32 * b1, nv) - min(32 * b0, nv))
        own_blocks = (b1 - b0) + (n_blocks - n_fine) / world
        rank_bytes = own_blocks * 4656 * 4 + 32 * own_verts
        roof = {"bound": "hbm", "kernel": "whole sharded apply, rank 0 share (level-0 solve dominates)",
                "achieved": rank_bytes / (ms_per_step * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": rank_bytes / (ms_per_step * 1e-3) / 1e9 / peak, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": int(rank_bytes), "traffic": None,
                "per_rank_fine_kernel_ms": per_rank_fine_ms,
                "aligned_cuts": g.aligned_cuts}
    # ---- e2e with the other ways a host caller can hand over r and z.  Runs last and guarded: a failure here leaves every
    # number above untouched.
    #   host_pull          MAS_OPT_HOST_PULL: a kernel pulls the page-locked residual instead of the copy engine (bit-identical)
    #   pageable           plain malloc'd buffers, what the C++ drop-in class sees from a std::vector caller
    #   pageable_registered   the same with MAS_OPT_REGISTER_HOST: the library page-locks the caller's buffers on first sight
    e2e_note = "mas_apply(MAS_MEM_HOST): pinned host r -> H2D -> apply graph -> D2H z, synchronous"
    e2e_modes = {"pinned_copy_engine": e2e_rate}
    e2e_staging = "pinned_copy_engine"
    if world == 1 and not args.lean:
        try:
            z_ce = z_h.clone()
            g.set_option(7, 1)
            pull_rate = time_e2e(e2e_steps)
            identical = bool(torch.equal(z_h, z_ce))
            g.set_option(7, 0)
            e2e_modes["host_pull"] = pull_rate
            e2e_modes["host_pull_bit_identical"] = identical
            if identical and pull_rate > e2e_rate:
                e2e_rate, e2e_staging = pull_rate, "host_pull (MAS_OPT_HOST_PULL=1)"
            r_pg, z_pg = r_np.copy(), np.zeros_like(r_np)
            e2e_modes["pageable"] = time_e2e(e2e_steps, z_pg, r_pg)
            g.set_option(9, 1)
            e2e_modes["pageable_registered"] = time_e2e(e2e_steps, z_pg, r_pg)
            e2e_modes["pageable_bit_identical"] = bool(np.array_equal(z_pg, z_ce.numpy()))
            g.set_option(9, 0)
        except Exception as exc:                      # noqa: BLE001
            e2e_modes["error"] = repr(exc)

    cfg["parallelism"] = ((f"morton-sharded x{world}, " + ("peer-memory exchange fused into the restriction kernel (NVLink)"
                                                            if p2p else "NCCL all-reduce of coarse residuals"))
                          if world > 1 else "single GPU")
    out = {
        "metric": METRIC, "value": units * 1e3 / ms_per_step, "unit": "applies/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak" if (weak or world == 1 and args.scaling == "weak") else "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "mesh": {"nv": nv, "levels": lv, "blocks": n_blocks, "mesh_applies_per_s": 1e3 / ms_per_step,
                 "exchange": "level-2 residuals over peer memory inside the apply graph" if p2p else ("NCCL all-reduce" if world > 1 else None)},
        "setup_ms": setup_ms, "setup_device_ms": setup_device_ms, "setup_rebuild_hierarchy_ms": setup_full_ms,
        "invert_variant": args.invert_variant,
        "e2e": {"value": units * e2e_rate, "unit": "applies/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "staging": e2e_staging, "applies_per_s_by_staging": e2e_modes,
                "note": ("every rank moves only its own vertices' r and z over its own PCIe link (page-locked buffers)" if world > 1 else e2e_note)},
        "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
        "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "parity": parity, "pcg": pcg,
        "strong_scaling_config4": strong4,
    }
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, help="index into BASELINE.json configs (default 2: 1M-vertex cloth)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1 with the default config: weak = one N x 1M-vertex cloth (default), strong = the 1M-vertex cloth split N ways")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lean", action="store_true", help="only the timed loop (for runs under ncu)")
    ap.add_argument("--variant", type=int, default=None, help="MAS_OPT_APPLY_VARIANT override (development sweeps)")
    ap.add_argument("--invert-variant", type=int, default=0, help="MAS_OPT_INVERT_VARIANT: 0 = tcgen05 tensor cores (default), 1 = FP32 CUDA cores")
    ap.add_argument("--nccl-exchange", action="store_true", help="N>1: use the NCCL all-reduce baseline instead of the peer-memory exchange")
    ap.add_argument("--cpu-pcg", action="store_true", help="also run the PCG solve on the host with the reference preconditioner")
    ap.add_argument("--proximity", action="store_true", help="config 1: folded sheet with the stencils of the proximity producer (collide.py)")
    ap.add_argument("--no-strong", action="store_true", help="skip the BASELINE config 4 (2048^2) strong-scaling measurement beside the headline")
    ap.add_argument("--no-arbiter", action="store_true", help="skip the FP64 arbiter of the in-run value parity (a few seconds per million vertices)")
    ap.add_argument("--cpu-leg-child", default=None, help=argparse.SUPPRESS)
    ap.add_argument("--cpu-leg-port", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_leg_child:
        cpu_leg_child(args.cpu_leg_child, args.steps, args.warmup, int(os.environ.get("MAS_BENCH_THREADS", host_threads())), args.cpu_leg_port)
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

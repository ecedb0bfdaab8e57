"""Sharded (multi-GPU) path on real hardware.  Runs LAST (file name): its multi-context and multi-process cases are the
ones most exposed to the box, and under `pytest -x` they must not hide the parity, PCG and limits suites.

  * two shards driven inside ONE process on one GPU (the exchange is done by adding the two exchange tensors): runs on the
    1-GPU box and exercises exactly the kernels/ranges a 2-GPU job runs;
  * the real thing over NCCL, one process per GPU, when the box has >= 2 GPUs (skipped otherwise).
Bar: merged z equals the single-device z to 1e-5 relative L2 (the only difference is the summation order of the FP64
coarse accumulators), every shard writes only its own vertices, structure identical on every rank."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

from helpers import rel_l2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_NAME = "preconditioner-for-cloth-and-deformable-body-simulation_b200"


def _mesh(synth, name):
    if name == "cloth96_collisions":
        m = synth.cloth(96, with_topology=True)
        return synth.add_collisions(m, m.nv // 16, m.nv // 16, m.nv // 8)
    if name == "cloth50_ragged":
        return synth.cloth(50)
    if name == "cloth256":
        return synth.cloth(256)
    if name == "cloth_rect512x256":
        return synth.cloth_rect(512, 256)          # 4 levels; level-1 blocks straddle the shard cuts at world = 3, 8
    raise KeyError(name)


@pytest.mark.parametrize("align", [1, 0])
@pytest.mark.parametrize("name,world", [("cloth96_collisions", 2), ("cloth50_ragged", 3), ("cloth256", 4),
                                        ("cloth_rect512x256", 8), ("cloth_rect512x256", 3)])
def test_shards_in_one_process_match_single_device(name, world, align, pkg, synth):
    import torch
    mesh = _mesh(synth, name)
    r = torch.from_numpy(synth.residual(mesh.nv)).cuda()
    single = pkg.SeSchwarzPreconditioner(0).setup_from_mesh(mesh, device_inputs=True)
    z1 = torch.empty_like(r)
    single.Preconditioning(z1, r)

    shards = [pkg.SeSchwarzPreconditioner(0, rank=k, world=world) for k in range(world)]
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    raw = lambda a: torch.from_numpy(np.frombuffer(np.ascontiguousarray(a).tobytes(), np.uint8).copy()).cuda()
    d = dict(pos=dev(mesh.positions), edges=dev(mesh.edges) if mesh.ne else None, faces=dev(mesh.faces) if mesh.nf else None,
             st=dev(mesh.nbr_starts), ix=dev(mesh.nbr_idx), diag=dev(mesh.diag), off=dev(mesh.offdiag),
             ef=raw(mesh.ef) if mesh.ef.size else None, ee=raw(mesh.ee) if mesh.ee.size else None, vf=raw(mesh.vf) if mesh.vf.size else None)
    for g in shards:
        g.set_option(4, align)                                      # MAS_OPT_ALIGN_CUTS: both exchange protocols
        g.m_positions, g.m_edges, g.m_faces, g.m_neighbours = d["pos"], d["edges"], d["faces"], (d["st"], d["ix"])
        g.AllocatePrecoditioner(mesh.nv, mesh.ne, mesh.nf)
        g.PreparePreconditioner(d["diag"], d["off"], d["st"], d["ef"], d["ee"], d["vf"], mesh.ef_total, mesh.ee_total, mesh.vf_total,
                                phase="begin")
    torch.cuda.synchronize()
    total = sum(g.exchange_tensor(0).clone() for g in shards)       # what one all-reduce does
    for g in shards:
        g.exchange_tensor(0).copy_(total)
        g.prepare_end()
    for g in shards:
        assert np.array_equal(g.going_next(), single.going_next())
        assert np.array_equal(g.level_size(), single.level_size())
    # both exchange protocols are exercised: level-2 residuals when every cut falls between level-1 banks, else level 1
    aligned = [g.aligned_cuts for g in shards]
    assert len(set(aligned)) == 1                      # every rank takes the same decision
    assert aligned[0] == bool(align), (name, aligned)  # these meshes all have an aligned bank within reach of every cut

    zs = [torch.full_like(r, float("nan")) for _ in shards]
    for g in shards:
        g.apply_begin(r)
    torch.cuda.synchronize()
    total = sum(g.exchange_tensor(1).clone() for g in shards)
    for g, z in zip(shards, zs):
        g.exchange_tensor(1).copy_(total)
        g.apply_end(z)
    torch.cuda.synchronize()
    s2o = torch.from_numpy(single.sorted_get_original().astype(np.int64)).cuda()
    merged = torch.zeros_like(r)
    covered = torch.zeros(mesh.nv, dtype=torch.int32, device="cuda")
    for g, z in zip(shards, zs):
        b, e = g.owned_fine_blocks
        own = s2o[min(32 * b, mesh.nv):min(32 * e, mesh.nv)]
        mask = torch.zeros(mesh.nv, dtype=torch.bool, device="cuda")
        mask[own] = True
        assert not torch.isnan(z[mask]).any() and torch.isnan(z[~mask]).all()    # a shard writes its own vertices only
        merged[mask] = z[mask]
        covered[mask] += 1
    assert bool((covered == 1).all())
    assert rel_l2(merged.cpu().numpy(), z1.cpu().numpy()) < 1e-5


@pytest.mark.parametrize("align", [1, 0])
@pytest.mark.parametrize("name,world", [("cloth96_collisions", 2), ("cloth256", 4), ("cloth_rect512x256", 3)])
def test_peer_memory_exchange_in_one_process(name, world, align, pkg, synth):
    """The production exchange (restriction kernel stores into every rank's arena + device-side flags) with all shards
    living in this process on one GPU, each on its own stream so that they really run concurrently and wait for each
    other on the device.  Results must equal the all-reduce protocol bit for bit, and repeat exactly.
    At most four shards here: every shard's graph has up to four branches that must all be co-scheduled with the other
    shards' spinning waits, which one GPU only guarantees while each stream has a hardware queue of its own
    (tests/conftest.py raises CUDA_DEVICE_MAX_CONNECTIONS to 32).  Eight shards are covered by real ranks
    (bench.py --gpus 8 checks the merged z against a single-GPU z in the run)."""
    import torch
    mesh = _mesh(synth, name)
    r = torch.from_numpy(synth.residual(mesh.nv)).cuda()
    streams = [torch.cuda.Stream() for _ in range(world)]
    shards = [pkg.SeSchwarzPreconditioner(0, rank=k, world=world, stream=streams[k]) for k in range(world)]
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    raw = lambda a: torch.from_numpy(np.frombuffer(np.ascontiguousarray(a).tobytes(), np.uint8).copy()).cuda()
    d = dict(pos=dev(mesh.positions), edges=dev(mesh.edges) if mesh.ne else None, faces=dev(mesh.faces) if mesh.nf else None,
             st=dev(mesh.nbr_starts), ix=dev(mesh.nbr_idx), diag=dev(mesh.diag), off=dev(mesh.offdiag),
             ef=raw(mesh.ef) if mesh.ef.size else None, ee=raw(mesh.ee) if mesh.ee.size else None, vf=raw(mesh.vf) if mesh.vf.size else None)
    torch.cuda.synchronize()
    for g in shards:
        g.set_option(4, align)
        g.m_positions, g.m_edges, g.m_faces, g.m_neighbours = d["pos"], d["edges"], d["faces"], (d["st"], d["ix"])
        g.AllocatePrecoditioner(mesh.nv, mesh.ne, mesh.nf)
    arenas = [g.peer_local() for g in shards]
    for g in shards:
        g.PreparePreconditioner(d["diag"], d["off"], d["st"], d["ef"], d["ee"], d["vf"], mesh.ef_total, mesh.ee_total, mesh.vf_total,
                                phase="begin")
    torch.cuda.synchronize()
    total = sum(g.exchange_tensor(0).clone() for g in shards)
    for g in shards:
        g.exchange_tensor(0).copy_(total)
    torch.cuda.synchronize()
    for g in shards:
        g.prepare_end()

    # reference: the all-reduce protocol on the same shards
    z_ref = [torch.zeros_like(r) for _ in shards]
    for g in shards:
        g.apply_begin(r)
    torch.cuda.synchronize()
    total = sum(g.exchange_tensor(1).clone() for g in shards)
    for g, z in zip(shards, z_ref):
        g.exchange_tensor(1).copy_(total)
    torch.cuda.synchronize()
    for g, z in zip(shards, z_ref):
        g.apply_end(z)
    torch.cuda.synchronize()

    for g in shards:
        g.peer_attach(pointers=arenas)
        if align == 0:
            g.set_option(11, 1)                            # MAS_OPT_STRICT_PUBLISH: system-scope fence before the flag stores
    zs = [torch.zeros_like(r) for _ in shards]             # same buffers every time: the apply graph is captured once per shard
    for rep in range(4):                                   # several applies: the double-buffered arenas and counters roll over
        for z in zs:
            z.zero_()
        torch.cuda.synchronize()
        for g, z in zip(shards, zs):
            g.Preconditioning(z, r)                        # enqueue only; the shards meet on the device
        torch.cuda.synchronize()
        for g, z, zr in zip(shards, zs, z_ref):
            assert g.peer_error == 0
            assert torch.equal(z, zr), rep
    merged = sum(zs)
    single = pkg.SeSchwarzPreconditioner(0).setup_from_mesh(mesh, device_inputs=True)
    z1 = torch.empty_like(r)
    single.Preconditioning(z1, r)
    torch.cuda.synchronize()
    assert rel_l2(merged.cpu().numpy(), z1.cpu().numpy()) < 1e-5


def _sharded_contexts(pkg, synth, mesh, world, streams):
    import torch
    shards = [pkg.SeSchwarzPreconditioner(0, rank=k, world=world, stream=streams[k]) for k in range(world)]
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    pos, st, ix, diag, off = dev(mesh.positions), dev(mesh.nbr_starts), dev(mesh.nbr_idx), dev(mesh.diag), dev(mesh.offdiag)
    torch.cuda.synchronize()
    for g in shards:
        g.m_positions, g.m_neighbours = pos, (st, ix)
        g.AllocatePrecoditioner(mesh.nv, 0, 0)
    arenas = [g.peer_local() for g in shards]
    for g in shards:
        g.PreparePreconditioner(diag, off, st, phase="begin")
    torch.cuda.synchronize()
    total = sum(g.exchange_tensor(0).clone() for g in shards)
    for g in shards:
        g.exchange_tensor(0).copy_(total)
    torch.cuda.synchronize()
    for g in shards:
        g.prepare_end()
        g.peer_attach(pointers=arenas)
    return shards


@pytest.mark.parametrize("name,world", [("cloth256", 2), ("cloth_rect512x256", 3)])
def test_host_pointer_apply_moves_only_owned_vertices(name, world, pkg, synth, pinned=True):
    """mas_apply(MAS_MEM_HOST) on a sharded context (what the C++ drop-in class calls): every shard runs in its own host
    thread, as a rank would.  A shard reads and writes ITS OWN vertices only — the other entries of the caller's z keep
    their values — and with page-locked buffers only those entries cross PCIe (2 x 16 B per owned vertex).
    (Pageable buffers take the copy engine, whose synchronous copies serialise the threads of ONE process against each other
    while a shard waits for its peer on the device; that branch is covered with real ranks in the 2-GPU test below.)"""
    import threading
    import torch
    mesh = _mesh(synth, name)
    r_np = synth.residual(mesh.nv)
    single = pkg.SeSchwarzPreconditioner(0).setup_from_mesh(mesh, device_inputs=True)
    z1 = np.zeros_like(r_np)
    single.Preconditioning(z1, r_np)
    streams = [torch.cuda.Stream() for _ in range(world)]
    shards = _sharded_contexts(pkg, synth, mesh, world, streams)
    if pinned:
        rs = [torch.from_numpy(r_np.copy()).pin_memory() for _ in range(world)]
        zs = [torch.full((mesh.nv, 4), 7.0).pin_memory() for _ in range(world)]
    else:
        rs = [r_np.copy() for _ in range(world)]
        zs = [np.full((mesh.nv, 4), 7.0, np.float32) for _ in range(world)]
    errors = []

    def run(k):
        try:
            for _ in range(3):
                shards[k].Preconditioning(zs[k], rs[k])
        except Exception as exc:                      # noqa: BLE001
            errors.append((k, repr(exc)))
    threads = [threading.Thread(target=run, args=(k,)) for k in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=60)
    assert not errors, errors
    s2o = single.sorted_get_original()
    merged = np.zeros_like(r_np)
    for k, g in enumerate(shards):
        z = zs[k].numpy() if pinned else zs[k]
        b, e = g.owned_fine_blocks
        own = s2o[min(32 * b, mesh.nv):min(32 * e, mesh.nv)]
        mask = np.zeros(mesh.nv, bool)
        mask[own] = True
        assert np.all(z[~mask] == 7.0)                          # foreign entries untouched
        merged[mask] = z[mask]
        owned_bytes = 16 * int(mask.sum())
        if pinned:
            assert g.get_int(16) == owned_bytes and g.get_int(17) == owned_bytes
        else:
            assert g.get_int(16) == 2 * 16 * mesh.nv and g.get_int(17) == 16 * mesh.nv
        assert g.peer_error == 0
    assert rel_l2(merged, z1) < 1e-5


def test_lost_peer_is_a_hard_error(pkg, synth):
    """A rank that never launches its apply: the device-side wait gives up after ~2 s, and instead of a silently wrong z the
    library reports it — mas_synchronize and every later mas_apply fail (sticky), peer_error reads 1."""
    import torch
    mesh = synth.cloth(256)
    streams = [torch.cuda.Stream() for _ in range(2)]
    shards = _sharded_contexts(pkg, synth, mesh, 2, streams)
    r = torch.from_numpy(synth.residual(mesh.nv)).cuda()
    z = torch.zeros_like(r)
    shards[0].Preconditioning(z, r)                   # shard 1 never shows up
    with pytest.raises(pkg.MasError, match="peer-memory exchange timed out"):
        shards[0].synchronize()
    with pytest.raises(pkg.MasError, match="peer-memory exchange timed out"):
        shards[0].Preconditioning(z, r)
    with pytest.raises(pkg.MasError, match="peer-memory exchange timed out"):
        shards[0].Preconditioning(np.zeros((mesh.nv, 4), np.float32), synth.residual(mesh.nv))
    assert shards[0].peer_error == 1


def _nccl_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        pkg = importlib.import_module(PKG_NAME)
        part = importlib.import_module(PKG_NAME + ".partition")
        mesh = pkg.synth.cloth(256)
        eng = pkg.SeSchwarzPreconditioner(rank, rank=rank, world=world, stream=torch.cuda.current_stream())
        drv = part.ShardedSchwarzPreconditioner(eng)
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        eng.m_positions, eng.m_neighbours = dev(mesh.positions), (dev(mesh.nbr_starts), dev(mesh.nbr_idx))
        drv.AllocatePrecoditioner(mesh.nv, 0, 0)
        drv.PreparePreconditioner(dev(mesh.diag), dev(mesh.offdiag), dev(mesh.nbr_starts))
        r = dev(pkg.synth.residual(mesh.nv))
        z = torch.zeros_like(r)
        drv.Preconditioning(z, r)
        torch.cuda.synchronize()
        z_allreduce = z.clone()
        # the same apply through the peer-memory exchange (IPC-mapped arenas): bit-identical shard results
        p2p = drv.attach_peers()
        same = True
        if p2p:
            z2 = torch.zeros_like(r)
            for _ in range(4):
                z2.zero_()
                drv.Preconditioning(z2, r)
                torch.cuda.synchronize()
                same = same and bool(torch.equal(z2, z_allreduce)) and eng.peer_error == 0
            # host-pointer apply (page-locked buffers): only the owned vertices cross PCIe, foreign entries stay untouched
            r_h = r.cpu().pin_memory()
            z_h = torch.zeros_like(r_h).pin_memory()
            drv.Preconditioning(z_h, r_h)
            same = same and bool(torch.equal(z_h.cuda(), z_allreduce)) and eng.get_int(16) < 16 * mesh.nv
            # pageable buffers: whole-array copies, foreign entries of the caller's z preserved
            r_p = r.cpu().numpy().copy()
            z_p = np.full((mesh.nv, 4), 7.0, np.float32)
            drv.Preconditioning(z_p, r_p)
            za = z_allreduce.cpu().numpy()
            own = za.any(axis=1)                      # the all-reduce run started from zeros: non-zero rows are owned
            same = same and np.array_equal(z_p[own], za[own]) and bool(np.all(z_p[~own] == 7.0)) and eng.get_int(17) == 16 * mesh.nv
        dist.all_reduce(z)                                        # disjoint shards, zeros elsewhere -> the full z
        torch.cuda.synchronize()
        if rank == 0:
            one = pkg.SeSchwarzPreconditioner(0).setup_from_mesh(mesh, device_inputs=True)
            z1 = torch.empty_like(r)
            one.Preconditioning(z1, r)
            torch.cuda.synchronize()
            q.put((float((z - z1)[:, :3].norm() / z1[:, :3].norm()), p2p, same))
        else:
            q.put((0.0, p2p, same))
    finally:
        dist.destroy_process_group()


def test_two_gpus_over_nccl_match_single_device():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = [ctx.Process(target=_nccl_worker, args=(k, 2, port, q)) for k in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    res = [q.get(timeout=5) for _ in range(2)]
    assert max(e for e, _, _ in res) < 1e-5
    assert all(p2p for _, p2p, _ in res), "CUDA IPC peer mapping unavailable on this box"
    assert all(same for _, _, same in res)

"""GPU tests at the edges of the input space that the earlier files do not reach: equal Morton codes (the tie rule),
repeated neighbour indices, and the largest mesh the five-level hierarchy holds.  Runs last (file name) so that a failure
here does not mask the parity, PCG and sharded suites under `pytest -x`."""
import numpy as np
import pytest

from helpers import assert_structure_equal, make_oracle
from test_gpu_parity import gpu_cls, test_structure_and_apply_vs_oracle as _structure_and_apply  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["stacked2x24", "stacked3x20", "cloth24_duplicate_edges"])
def test_ties_and_duplicate_neighbours_vs_oracle(name, gpu_cls, synth, oracle_lib, monkeypatch):
    """stacked*: coincident sheets = groups of equal Morton codes; the stable radix sort over an iota payload must give the
    oracle's (code, original index) order, and everything downstream of it bit for bit.  duplicate_edges: neighbour lists
    with repeated indices, each with its own 3x3 block (they add up, cpp:1292-1298)."""
    import test_gpu_parity as tp
    mesh = {"stacked2x24": lambda: synth.stacked_cloth(24, 2), "stacked3x20": lambda: synth.stacked_cloth(20, 3),
            "cloth24_duplicate_edges": lambda: synth.cloth_with_duplicate_edges(24)}[name]
    monkeypatch.setattr(tp, "_cases", lambda s: {name: mesh})
    _structure_and_apply(name, gpu_cls, synth, oracle_lib)


@pytest.mark.parametrize("args", [(1500, 5, 3), (4000, 7, 4), (900, 3, 5), (2500, 10, 6)])
def test_irregular_point_clouds_vs_oracle(args, gpu_cls, synth, oracle_lib, monkeypatch):
    """Random 3-D points joined to their k nearest neighbours: degrees from 3 to 19, Morton banks that cut through the
    connectivity anywhere.  tests/test_oracle_vs_reference.py pins the oracle to the compiled reference on the same four
    meshes; here the CUDA path has to match the oracle (integers bit for bit, inverses and z within the parity bars)."""
    import test_gpu_parity as tp
    name = "cloud%d_k%d" % args[:2]
    monkeypatch.setattr(tp, "_cases", lambda s: {name: lambda: synth.random_cloud(*args)})
    _structure_and_apply(name, gpu_cls, synth, oracle_lib)


@pytest.mark.parametrize("n", [64, 96])
def test_rippled_cloth_with_fragmented_banks_vs_oracle(n, gpu_cls, synth, oracle_lib, monkeypatch):
    """A 1e-4 out-of-plane ripple gets the full weight of the z bits of the Morton code (per-axis normalisation, cpp:225):
    banks hold several components each, level 1 is 4-5 times larger than on the flat sheet and the top level keeps several
    nodes.  The reference overruns its fixed allocation on such input from 512^2 on (Q6); buffers here follow the counts."""
    import test_gpu_parity as tp
    name = f"rippled{n}"
    monkeypatch.setattr(tp, "_cases", lambda s: {name: lambda: synth.rippled_cloth(n)})
    _structure_and_apply(name, gpu_cls, synth, oracle_lib)


def test_edge_free_particles_grow_every_level(gpu_cls, synth, oracle_lib):
    """6000 free particles: every level keeps 6000 one-vertex clusters (the node arrays have to grow past their first
    guess of nv + nv/8 + 4096) and every domain matrix is the identity: structure equals the oracle's, z = 3 r exactly."""
    mesh = synth.dust(6000)
    g = gpu_cls(0).setup_from_mesh(mesh)
    o = make_oracle(oracle_lib, mesh)
    assert_structure_equal(g, o, mesh.nv)
    assert g.level_size().tolist() == [[0, 0], [6000, 6016], [6000, 12032], [6000, 18048]]
    r = synth.residual(mesh.nv)
    z = np.full_like(r, 5.0)
    g.Preconditioning(z, r)
    assert np.array_equal(z[:, :3], np.float32(3) * r[:, :3]) and np.all(z[:, 3] == 0)
    for b in (0, 187, 188, 400, g.num_blocks - 1):
        assert np.array_equal(g.dense_inverse(b), np.eye(96, dtype=np.float32))


def test_one_vertex_over_the_five_level_limit_is_refused(gpu_cls, pkg):
    """32^5 + 1 vertices need a sixth level, which the Int4 ancestor table (SeSchwarzPreconditioner.h:96) cannot hold: the
    reference would overrun it; the C ABI returns MAS_ERR_UNSUPPORTED before touching any input."""
    g = gpu_cls(0)
    g.m_positions = np.zeros((1, 4), np.float32)
    g.m_neighbours = (np.zeros(2, np.int32), np.zeros(1, np.int32))
    with pytest.raises(pkg.MasError, match="more than 5 levels"):
        g.AllocatePrecoditioner(32 ** 5 + 1, 0, 0)


def test_maximum_size_33_million_vertices(gpu_cls, synth):
    """The largest supported mesh: 8192 x 4096 cloth = 33,554,432 vertices = exactly five full levels, 1,082,401 domains,
    20.2 GB of packed inverses (64 GB of HBM in use with the caller's Hessian).  Hierarchy sizes are exact; M^-1 is checked
    through its size-independent properties (tools/max_size_check.py prints the timings: setup 151 ms, apply 3.28 ms)."""
    import torch
    dev = torch.device("cuda:0")
    free, _ = torch.cuda.mem_get_info()
    if free < 100e9:
        pytest.skip("needs ~70 GB of free HBM")
    mesh = synth.cloth_rect_device(8192, 4096, dev)
    nv = mesh.nv
    g = gpu_cls(0)
    g.m_positions, g.m_neighbours = mesh.positions, (mesh.nbr_starts, mesh.nbr_idx)
    g.AllocatePrecoditioner(nv, 0, 0)
    g.PreparePreconditioner(mesh.diag, mesh.offdiag, mesh.nbr_starts)
    assert g.num_level == 5 and g.num_blocks == 1082401
    assert g.level_size().tolist() == [[0, 0], [1048576, 33554432], [32768, 34603008], [1024, 34635776], [32, 34636800],
                                       [1, 34636832]]
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
    assert bool(torch.isfinite(z1).all()) and bool((z1[:, 3] == 0).all())
    assert abs(d(r1, z2) - d(r2, z1)) <= 1e-6 * abs(d(r1, z1))
    assert d(r1, z1) > 0 and d(r2, z2) > 0
    lin = 2.0 * z1 - 0.5 * z2
    assert float((z12 - lin)[:, :3].norm() / lin[:, :3].norm()) < 1e-5
    assert torch.equal(z1, z1b)
    g.close()
    del mesh, r1, r2, z1, z2, z12, z1b, lin
    torch.cuda.empty_cache()


def test_host_pull_staging_is_bit_identical(gpu_cls, synth):
    """MAS_OPT_HOST_PULL: a page-locked residual is pulled by a kernel through its device mapping instead of the copy
    engine; pageable buffers keep the copy engine.  Same z, bit for bit, in every combination."""
    import torch
    mesh = synth.cloth(96)
    g = gpu_cls(0).setup_from_mesh(mesh)
    r_np = synth.residual(mesh.nv)
    r_pin = torch.from_numpy(r_np.copy()).pin_memory()
    z_ce, z_pull = torch.zeros_like(r_pin).pin_memory(), torch.full_like(r_pin, 7.0).pin_memory()
    g.Preconditioning(z_ce, r_pin)
    g.set_option(7, 1)
    g.Preconditioning(z_pull, r_pin)
    assert torch.equal(z_ce, z_pull)
    z_pageable = np.full_like(r_np, 3.0)
    g.Preconditioning(z_pageable, r_np)                   # pageable r: copy engine even with the option on
    assert np.array_equal(z_pageable, z_ce.numpy())
    g.set_option(7, 2)                                    # auto: six sampled applies, then one staging is kept
    assert g.get_int(15) == -1
    for _ in range(7):
        z_auto = torch.full_like(r_pin, -1.0).pin_memory()
        g.Preconditioning(z_auto, r_pin)
        assert torch.equal(z_auto, z_ce)
    assert g.get_int(15) in (0, 1)
    g.set_option(7, 0)


import os  # noqa: E402


@pytest.mark.parametrize("name", ["cloth64", "cloth96_collisions", "tet16x16x8", "cloth200_stiff", "chain100_fragmented_banks"])
def test_cuda_core_inversion_holds_the_parity_bar(name, gpu_cls, synth, oracle_lib):
    """MAS_OPT_INVERT_VARIANT = 1: the FP32 CUDA-core inversion (the reference's elimination regrouped by tiles) instead of the
    default tensor-core kernel.  Both have to pass the very same parity test (structure bit-exact, inverses and z within the
    FP64-arbitrated bars); every other GPU test runs the tensor-core default."""
    def with_variant(device):
        g = gpu_cls(device)
        g.set_option(8, 1)
        return g
    _structure_and_apply(name, with_variant, synth, oracle_lib)


def test_inversion_kernels_agree_on_every_block(gpu_cls, synth):
    """Tensor-core (3xTF32 block Gauss-Jordan) against CUDA-core (FP32 LDL^T) inverses of the same setup, every block of a
    stiff cloth with collisions: two different algorithms, each within rounding of the true inverse."""
    m = synth.cloth(96, k=1e4, with_topology=True)
    mesh = synth.add_collisions(m, m.nv // 16, m.nv // 16, m.nv // 8)
    a = gpu_cls(0).setup_from_mesh(mesh)
    b = gpu_cls(0)
    b.set_option(8, 1)
    b.setup_from_mesh(mesh)
    worst = 0.0
    for blk in range(a.num_blocks):
        ia, ib = a.dense_inverse(blk), b.dense_inverse(blk)
        assert np.array_equal(ia, ia.T)
        worst = max(worst, float(np.abs(ia - ib).max() / np.abs(ib).max()))
    assert worst < 2e-5, worst


def test_register_host_option_keeps_results(gpu_cls, synth):
    """MAS_OPT_REGISTER_HOST page-locks the caller's pageable r / z where they lie; results are unchanged, with and without
    the kernel pull, and the ranges are released when the option is cleared."""
    mesh = synth.cloth(96)
    g = gpu_cls(0).setup_from_mesh(mesh)
    r = synth.residual(mesh.nv)
    z0 = np.zeros_like(r)
    g.Preconditioning(z0, r)
    z = np.zeros_like(r)                  # the contract: registered buffers stay allocated until the option is cleared
    g.set_option(9, 1)
    for pull in (0, 1):
        g.set_option(7, pull)
        for _ in range(3):
            z[:] = 9.0
            g.Preconditioning(z, r)
            assert np.array_equal(z, z0)
    g.set_option(9, 0)
    g.set_option(7, 0)
    z[:] = 9.0
    g.Preconditioning(z, r)
    assert np.array_equal(z, z0)


def test_cached_hierarchy_gives_identical_setup(gpu_cls, synth):
    """MAS_OPT_CACHE_HIERARCHY (default on): collision-free prepares keep the clustering; a prepare with stencils rebuilds it."""
    m = synth.cloth(96, with_topology=True)
    coll = synth.add_collisions(synth.cloth(96, with_topology=True), m.nv // 16, m.nv // 16, m.nv // 8)   # (adds to its argument)
    stiff = synth.cloth(96, k=5000.0)
    r = synth.residual(m.nv)
    assert m.ef_total == 0 and coll.ef_total > 0

    def run(cache):
        g = gpu_cls(0)
        g.set_option(10, cache)
        g.setup_from_mesh(m)
        out = []
        for mesh in (stiff, coll, m, stiff):
            g.PreparePreconditioner(mesh.diag, mesh.offdiag, mesh.nbr_starts, mesh.ef, mesh.ee, mesh.vf, mesh.ef_total, mesh.ee_total,
                                    mesh.vf_total)
            z = np.zeros_like(r)
            g.Preconditioning(z, r)
            out.append((g.going_next().copy(), g.level_size().copy(), z, g.prepare_launches))
        return out
    a, b = run(0), run(1)
    for k, ((ga, la, za, na), (gb, lb, zb, nb)) in enumerate(zip(a, b)):
        assert np.array_equal(ga, gb) and np.array_equal(la, lb)
        # the coarse Galerkin sums are FP64 atomics (order-dependent in the last bit, Q7): z agrees to rounding, not bit for bit
        assert np.abs(za - zb).max() <= 1e-6 * np.abs(za).max()
    na, nb = [x[3] for x in a], [x[3] for x in b]
    assert nb[0] < na[0] and nb[3] < na[3], (na, nb)   # fewer launches when the clustering is kept
    assert nb[2] == na[2], (na, nb)                    # first collision-free prepare after one with stencils rebuilds


@pytest.mark.parametrize("n", [64, 192, 512])
def test_graph_and_plain_launch_sequences_agree_bit_for_bit(n, gpu_cls, synth):
    """The captured apply graph (two branches; on small meshes the shortened coarse chain: top walk from level 1, ancestor walk
    instead of prolong_sum) against the strictly sequential un-captured launches: same arithmetic in the same order, so z must
    not differ by a bit."""
    import torch
    mesh = synth.cloth(n)
    g = gpu_cls(0).setup_from_mesh(mesh, device_inputs=True)
    r = torch.from_numpy(synth.residual(mesh.nv)).cuda()
    z0, z1 = torch.empty_like(r), torch.empty_like(r)
    g.Preconditioning(z0, r)
    g.set_option(2, 0)                    # MAS_OPT_USE_GRAPH = 0
    g.set_option(1, 0)                    # MAS_OPT_APPLY_VARIANT = 0: no concurrent head
    g.Preconditioning(z1, r)
    torch.cuda.synchronize()
    assert torch.equal(z0, z1)


def test_context_lifecycle_returns_all_device_memory(gpu_cls, synth, pkg):
    """Thirty create / allocate / prepare / apply / solve / destroy cycles on meshes of changing size, some with collision
    stencils, and an attempt to re-allocate an object for another mesh (refused): the free device memory after the last destroy equals the free memory
    before the first create (the library owns no allocator cache), the persisting-L2 carve-out of mas_pcg_solve is released,
    and z of the last cycle equals z of the first cycle on the same mesh (to the last bits: the coarse Galerkin sums are FP64
    atomics whose order varies from setup to setup)."""
    import gc
    import torch
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    meshes = [synth.cloth(40), synth.cloth(96), synth.tet_cube(12, 12, 6)]
    m = synth.cloth(64, with_topology=True)
    meshes.append(synth.add_collisions(m, 40, 40, 80))
    first = None
    free0 = None
    for cycle in range(31):
        mesh = meshes[cycle % len(meshes)] if cycle < 30 else meshes[0]
        g = gpu_cls(0).setup_from_mesh(mesh)
        r = synth.residual(mesh.nv)
        z = np.zeros_like(r)
        g.Preconditioning(z, r)
        if cycle % 3 == 0:
            res = pkg.pcg_solve(g, mesh.diag, mesh.offdiag, mesh.nbr_starts, mesh.nbr_idx, r, max_iter=12)
            assert res.iterations > 0
        if cycle % 4 == 1:                                     # a mesh of other sizes on the same object is refused, not run
            other = meshes[(cycle + 1) % len(meshes)]
            with pytest.raises(pkg.MasError, match="one mesh per object"):
                g.setup_from_mesh(other)
            z2 = np.zeros_like(r)
            g.Preconditioning(z2, r)                            # ... and the object still serves its own mesh
            assert np.array_equal(z2, z)
        if cycle == 0:
            first = z.copy()
        g.close()
        del g
        gc.collect()
        torch.cuda.synchronize()
        if cycle == 0:
            free0 = torch.cuda.mem_get_info()[0]                # after one full cycle: CUDA's own lazily created state is in place
    assert np.linalg.norm(z - first) <= 1e-6 * np.linalg.norm(first)
    free1 = torch.cuda.mem_get_info()[0]
    assert free0 - free1 <= 4 << 20, (free0, free1)             # nothing accumulates (allow 4 MiB for allocator granularity)


def test_stencil_counts_grow_and_shrink_between_prepares(gpu_cls, synth):
    """One object, PreparePreconditioner with few, then eight times as many, then few collision stencils again (the stencil
    buffers grow on the way): every z equals that of a fresh object prepared once with the same input."""
    def coll(frac):
        m = synth.cloth(96, with_topology=True)
        return synth.add_collisions(m, m.nv // (4 * frac), m.nv // (4 * frac), m.nv // (2 * frac))
    few, many = coll(8), coll(1)
    assert many.ef_total >= 8 * few.ef_total > 0
    r = synth.residual(few.nv)

    def prepare_apply(g, mesh):
        g.PreparePreconditioner(mesh.diag, mesh.offdiag, mesh.nbr_starts, mesh.ef, mesh.ee, mesh.vf, mesh.ef_total, mesh.ee_total,
                                mesh.vf_total)
        z = np.zeros_like(r)
        g.Preconditioning(z, r)
        return z, g.stencil_num, g.level_size().copy()
    g = gpu_cls(0).setup_from_mesh(few)
    for mesh in (few, many, few, many):
        z, n, levels = prepare_apply(g, mesh)
        z1, n1, levels1 = prepare_apply(gpu_cls(0).setup_from_mesh(mesh), mesh)
        # (the reference keeps only the stencils that pass its own filters: fewer than the totals handed in)
        assert 0 < n == n1 <= mesh.ef_total + mesh.ee_total + mesh.vf_total and np.array_equal(levels, levels1)
        assert np.abs(z - z1).max() <= 1e-6 * np.abs(z1).max()


@pytest.mark.parametrize("n", [64, 512])
def test_three_thousand_graph_replays_are_bit_identical(n, gpu_cls, synth):
    """Soak: 3,000 back-to-back applies of the captured graph (coarse chain, head, tail and coarse addition overlap inside it),
    alternating between two residuals and rotating through seven output buffers: every z equals the first z of its residual bit for bit, and
    nothing of one apply leaks into the next."""
    import torch
    mesh = synth.cloth_rect_device(n, n, torch.device("cuda:0"))
    g = gpu_cls(0)
    g.m_positions, g.m_neighbours = mesh.positions, (mesh.nbr_starts, mesh.nbr_idx)
    g.AllocatePrecoditioner(mesh.nv, 0, 0)
    g.PreparePreconditioner(mesh.diag, mesh.offdiag, mesh.nbr_starts)
    ra = torch.from_numpy(synth.residual(mesh.nv)).cuda()
    rb = torch.flip(ra, dims=[0]).contiguous() * 3.0
    za, zb = torch.empty_like(ra), torch.empty_like(ra)
    g.Preconditioning(za, ra)
    g.Preconditioning(zb, rb)
    torch.cuda.synchronize()
    first_a, first_b = za.clone(), zb.clone()
    assert not torch.equal(first_a, first_b)
    bad = torch.zeros((), dtype=torch.int64, device="cuda")
    # six output buffers for the first residual, one for the second: seven (residual, z) pointer pairs in rotation, more than the
    # context keeps instantiated graphs for (four; the least recently used one is re-captured)
    zs = [za] + [torch.empty_like(ra) for _ in range(5)]
    for k in range(1500):
        z = zs[k % len(zs)]
        z.fill_(float("nan"))
        g.Preconditioning(z, ra)
        bad += (z != first_a).any().long()
        zb.fill_(float("nan"))
        g.Preconditioning(zb, rb)
        bad += (zb != first_b).any().long()
    torch.cuda.synchronize()
    assert int(bad) == 0


def test_two_contexts_on_one_gpu_from_two_host_threads(gpu_cls, synth, pkg):
    """The library keeps no state outside its handles: two objects driven at the same time from two host threads, each on a
    stream of its own (different meshes, setup + 200 applies + a PCG solve), give what each gives alone."""
    import threading
    import torch
    meshes = [synth.cloth(96), synth.tet_cube(16, 16, 8)]

    def run(mesh, stream, out, key):
        try:
            with torch.cuda.stream(stream):
                g = gpu_cls(0, stream=stream).setup_from_mesh(mesh, device_inputs=True)
                r = torch.from_numpy(synth.residual(mesh.nv)).cuda()
                z = torch.empty_like(r)
                for _ in range(200):
                    g.Preconditioning(z, r)
                d = g._dev_inputs
                res = pkg.pcg_solve(g, d[0], d[1], d[2], torch.from_numpy(mesh.nbr_idx).cuda(), r, rel_tol=1e-4)
                stream.synchronize()
                out[key] = (z.cpu().numpy(), res.iterations, res.x.cpu().numpy())
                g.close()
        except Exception as e:                                   # surfaced by the assertion below
            out[key] = e
    alone, together = {}, {}
    for i, mesh in enumerate(meshes):
        run(mesh, torch.cuda.Stream(), alone, i)
    threads = [threading.Thread(target=run, args=(mesh, torch.cuda.Stream(), together, i)) for i, mesh in enumerate(meshes)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    for i in range(len(meshes)):
        assert not isinstance(alone[i], Exception), alone[i]
        assert not isinstance(together[i], Exception), together[i]
        # (setup sums its coarse Galerkin terms with FP64 atomics: two setups of one mesh agree to rounding, not bit for bit)
        assert np.abs(together[i][0] - alone[i][0]).max() <= 1e-6 * np.abs(alone[i][0]).max()
        assert abs(together[i][1] - alone[i][1]) <= 1
        assert np.linalg.norm(together[i][2] - alone[i][2]) <= 1e-3 * np.linalg.norm(alone[i][2])


def test_phase_split_and_graph_applies_do_not_share_state(gpu_cls, synth):
    """mas_apply (captured graph, cached per pointer pair) and the phase-split pair mas_apply_begin / mas_apply_end on one
    object, with pointers chosen so that a shared "last residual" field would replay a stale graph."""
    import torch
    mesh = synth.cloth(96)
    g = gpu_cls(0).setup_from_mesh(mesh)
    r1 = torch.from_numpy(synth.residual(mesh.nv)).cuda()
    r2 = (r1 * -2.0).contiguous()
    z1, z2, z3 = torch.empty_like(r1), torch.empty_like(r1), torch.empty_like(r1)
    g.Preconditioning(z1, r1)            # graph for (r1, z1)
    g.apply_begin(r2)
    g.apply_end(z2)                      # phase split with r2: z2 = M^-1 r2
    g.Preconditioning(z1, r2)            # new pair (r2, z1): must not replay the (r1, z1) graph
    g.Preconditioning(z3, r1)
    torch.cuda.synchronize()
    assert torch.equal(z1, z2)
    assert torch.allclose(z3 * -2.0, z2, rtol=0, atol=1e-5 * float(z2.abs().max()))

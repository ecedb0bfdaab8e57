"""World-size-2 test of the multi-GPU host logic on CPU (gloo): the sharding, what is exchanged, and that the sum of
the shards is the single-device result.  The per-rank CUDA engine is replaced by a small numpy engine that follows the
same begin -> exchange -> end protocol (and the same ownership rules as csrc/mas_assemble.cu / mas_apply.cu); the
driver under test (partition.ShardedSchwarzPreconditioner) is the production one.  Checked against the FP64 oracle."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_NAME = "preconditioner-for-cloth-and-deformable-body-simulation_b200"


class NumpyShardEngine:
    """One shard, FP64, structure taken from the oracle (the hierarchy is replicated on every rank in production too)."""

    def __init__(self, oracle, mesh, rank, world, part):
        self.o, self.mesh = oracle, mesh
        self.nv = mesh.nv
        self.L = oracle.num_level
        self.tc = oracle.total_clusters
        self.nVC = (self.nv + 31) // 32 * 32
        self.gn = np.asarray(oracle.going_next())[:self.tc]
        self.s2o = oracle.sorted_get_original()
        self.adj_s, self.adj_i = oracle.sorted_adjacency()
        self.vb, self.ve = part.owned_vertex_range(self.nv, rank, world)
        self.ncb = (self.tc - self.nVC) // 32
        self.acc = np.zeros(self.ncb * 96 * 96 + (self.tc - self.nVC) * 9, np.float64)
        self.R = np.zeros((self.tc - self.nVC) * 4, np.float32)

    def AllocatePrecoditioner(self, *a):  # noqa: N802
        pass

    def exchange_tensor(self, which):
        import torch
        return torch.from_numpy(self.acc if which == 0 else self.R)

    def _chain(self, v):
        out = [v]
        for _ in range(self.L - 1):
            out.append(int(self.gn[out[-1]]))
        return out

    def PreparePreconditioner(self, diag, off, ranges, *a, phase=None):  # noqa: N802
        assert phase == "begin"
        self.acc[:] = 0
        dense = self.acc[:self.ncb * 9216].reshape(self.ncb, 96, 96)
        carry = self.acc[self.ncb * 9216:].reshape(-1, 3, 3)       # per coarse node: everything added to its own diagonal block
        for v in range(self.vb, self.ve):
            ov = int(self.s2o[v])
            cv = self._chain(v)
            D = diag[ov].reshape(3, 3).T.astype(np.float64).copy()   # column-major SeMatrix3f -> (i,j)
            for k, e in enumerate(range(self.adj_s[v], self.adj_s[v + 1])):
                u = int(self.adj_i[e])
                M = off[ranges[ov] + k].reshape(3, 3).T.astype(np.float64)
                cu = self._chain(u)
                lvl = next((l for l in range(self.L) if cv[l] // 32 == cu[l] // 32), None)
                if lvl is None:
                    continue                                        # cpp:1288-1291
                if lvl == 0:
                    D += M                                          # folded into the diagonal that moves upward (cpp:1297-1298)
                    continue
                a, b = cv[lvl] - self.nVC, cu[lvl] - self.nVC
                dense[a // 32, 3 * (a % 32):3 * (a % 32) + 3, 3 * (b % 32):3 * (b % 32) + 3] += M
                for l in range(lvl + 1, self.L):
                    carry[cv[l] - self.nVC] += M
            for l in range(1, self.L):
                carry[cv[l] - self.nVC] += D

    def prepare_end(self):
        dense = self.acc[:self.ncb * 9216].reshape(self.ncb, 96, 96).copy()
        carry = self.acc[self.ncb * 9216:].reshape(-1, 3, 3)
        self.cinv = np.zeros_like(dense)
        for b in range(self.ncb):
            for n in range(32):
                dense[b, 3 * n:3 * n + 3, 3 * n:3 * n + 3] += carry[32 * b + n]
                if dense[b, 3 * n, 3 * n] == 0.0:                   # padding node -> identity (cpp:1365-1368)
                    dense[b, 3 * n:3 * n + 3, 3 * n:3 * n + 3] = np.eye(3)
            self.cinv[b] = np.linalg.inv(dense[b])

    def apply_begin(self, r):
        self.r = r                                                   # like mas_apply_begin, which keeps the pointer
        R = self.R.reshape(-1, 4)
        R[:] = 0
        for v in range(self.vb, self.ve):
            R[self.gn[v] - self.nVC, :3] += r[self.s2o[v], :3]

    def apply_end(self, z):
        r = self.r
        R = self.R.reshape(-1, 4)[:, :3].astype(np.float64)
        ls = np.asarray(self.o.level_size())
        for l in range(1, self.L - 1):                               # restrict l -> l+1
            for c in range(int(ls[l][1]), int(ls[l][1]) + int(ls[l][0])):
                R[self.gn[c] - self.nVC] += R[c - self.nVC]
        Z = np.einsum("bij,bj->bi", self.cinv, R.reshape(self.ncb, 96)).reshape(-1, 3)
        top = min(self.L, 4)
        for b in range(self.vb // 32, (self.ve + 31) // 32):
            inv = self.o.dense_inverse(b).astype(np.float64)
            x = np.zeros(96)
            n = min(32, self.nv - 32 * b)
            x[:3 * n] = r[self.s2o[32 * b:32 * b + n], :3].ravel()
            y = (inv @ x).reshape(32, 3)
            for i in range(n):
                v = 32 * b + i
                acc, node = y[i].copy(), v
                for _ in range(1, top):
                    node = int(self.gn[node])
                    acc += Z[node - self.nVC]
                z[self.s2o[v], :3] = acc
                z[self.s2o[v], 3] = 0


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pkg = importlib.import_module(PKG_NAME)
        part = importlib.import_module(PKG_NAME + ".partition")
        from oracle import oracle_binding as ob
        mesh = pkg.synth.cloth(40)
        o = ob.OraclePreconditioner("d")
        o.allocate(mesh)
        o.prepare()
        eng = NumpyShardEngine(o, mesh, rank, world, part)
        drv = part.ShardedSchwarzPreconditioner(eng)
        assert (drv.rank, drv.world) == (rank, world)
        drv.AllocatePrecoditioner(mesh.nv, 0, 0)
        drv.PreparePreconditioner(mesh.diag.reshape(-1, 9), mesh.offdiag.reshape(-1, 9), mesh.nbr_starts)
        # after the exchange every rank holds the FULL coarse Galerkin blocks: their inverses equal the oracle's
        nfb = eng.nVC // 32
        worst = max(np.abs(eng.cinv[b] - o.dense_inverse(nfb + b)).max() / np.abs(o.dense_inverse(nfb + b)).max() for b in range(eng.ncb))
        r = pkg.synth.residual(mesh.nv)
        z = np.full_like(r, np.nan)
        drv.Preconditioning(z, r)
        own = np.zeros(mesh.nv, bool)
        own[eng.s2o[eng.vb:eng.ve]] = True
        assert np.isnan(z[~own]).all() and not np.isnan(z[own]).any()      # a shard writes only its own vertices
        zt = torch.from_numpy(np.nan_to_num(z))
        full = drv.gather_z(zt, torch.from_numpy(own)).numpy()
        z_ref = o.apply(r)
        err = np.linalg.norm(full[:, :3] - z_ref[:, :3]) / np.linalg.norm(z_ref[:, :3])
        q.put((rank, float(worst), float(err), (eng.vb, eng.ve)))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_bank_ranges_tile_the_mesh(pkg):
    part = importlib.import_module(PKG_NAME + ".partition")
    for n in (1, 7, 50, 32768, 131072):
        for world in (1, 2, 3, 4, 8):
            ranges = [part.fine_bank_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in ranges]
            assert max(sizes) - min(sizes) <= 1
    assert part.owned_vertex_range(1600, 1, 2) == (800, 1600)
    assert part.owned_vertex_range(49, 1, 2) == (32, 49)
    with pytest.raises(ValueError):
        part.fine_bank_range(10, 2, 2)


def test_two_shards_over_gloo_equal_single_device(oracle_lib):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    res = sorted(q.get(timeout=5) for _ in range(2))
    assert res[0][3] == (0, 800) and res[1][3] == (800, 1600)
    for _, worst_inv, err, _ in res:
        assert worst_inv < 1e-7      # summed Galerkin accumulators reproduce the single-device coarse blocks
        assert err < 2e-6            # sum of the shards equals the FP64 oracle apply (the exchanged residuals are FP32, as in production)

"""Multi-GPU host side: one process per GPU, Morton-contiguous shards of the 32-node fine domains.

The reference is single-process (OpenMP only); this layer is new.  What shards and what is exchanged:

  * fine domains (97 % of the bytes) are split into `world` contiguous ranges of banks in Morton order; rank g
    assembles, inverts and applies only its own banks and produces z only for its own vertices.  The CUDA engine moves
    the cuts (during PreparePreconditioner) to banks where the running level-1 id is a multiple of 32, so that no level-1
    bank straddles two shards: query `engine.owned_fine_blocks` after the first prepare;
  * level-1 nodes are owned by exactly one fine bank (clustering is per bank, cpp:565-740), so the restriction
    r -> R_1 is local, and with aligned cuts so are the level-1 solves and the restriction R_1 -> R_2; every rank then
    needs the complete level-2 residual (nv/1024 nodes): the one exchange per apply, `exchange(1)`;
  * setup has ONE sum over ranks of the coarse Galerkin accumulators (FP64): `exchange(0)`;
  * levels >= 2 are solved redundantly on every rank, which removes any exchange of z;
  * production path for the apply exchange: `attach_peers()` maps every rank's exchange arena into every other rank
    (CUDA IPC handles travel over torch.distributed once); from then on a kernel inside the apply graph publishes the
    rank's residuals in its own arena, raises a flag in every peer, waits for the peers' flags on the device and pulls
    their slices over NVLink: Preconditioning() is a single graph launch per rank — no NCCL call and no host sync per
    apply.  Without attached peers the exchange is one `all_reduce` between apply_begin and apply_end (the baseline).

`ShardedSchwarzPreconditioner` drives a per-rank engine through begin -> exchange -> end.  The engine is
`SeSchwarzPreconditioner` (CUDA, exchange buffers are device tensors, NCCL) in production; the CPU tests inject a
numpy engine and run the same driver over gloo.
"""
from __future__ import annotations

from typing import Optional, Tuple


def fine_bank_range(n_fine_banks: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of the fine banks rank `rank` owns under the EVEN split (the numpy engine of the CPU tests, and the CUDA
    engine before its first prepare or with MAS_OPT_ALIGN_CUTS = 0)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return n_fine_banks * rank // world, n_fine_banks * (rank + 1) // world


def owned_vertex_range(nv: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) in SORTED (Morton) vertex ids."""
    n_banks = (nv + 31) // 32
    b, e = fine_bank_range(n_banks, rank, world)
    return min(b * 32, nv), min(e * 32, nv)


class ShardedSchwarzPreconditioner:
    """The reference's three calls (h:55-63) over `world` shards.

    engine must provide: AllocatePrecoditioner, PreparePreconditioner(..., phase="begin"), prepare_end(),
    apply_begin(r), apply_end(z), exchange_tensor(which) -> torch tensor aliasing the exchange buffer."""

    def __init__(self, engine, group=None):
        import torch.distributed as dist
        self.engine = engine
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._apply_exchange = None
        self.p2p = False

    def attach_peers(self) -> bool:
        """Exchange the IPC handles of the per-rank arenas and attach them (call after AllocatePrecoditioner).
        Returns False (and stays on the all-reduce path) if peer mapping is not available on this machine."""
        if self.world < 2:
            return False
        handles = [None] * self.world
        try:
            mine = self.engine.peer_export()
        except Exception:
            mine = None
        self.dist.all_gather_object(handles, mine, group=self.group)
        ok = all(h is not None for h in handles)
        if ok:
            try:
                self.engine.peer_attach(handles=handles)
            except Exception:
                ok = False
        flags = [None] * self.world
        self.dist.all_gather_object(flags, ok, group=self.group)
        self.p2p = all(flags)
        return self.p2p

    def _sum_over_ranks(self, tensor):
        if self.world > 1 and tensor.numel():
            self.dist.all_reduce(tensor, op=self.dist.ReduceOp.SUM, group=self.group)

    def AllocatePrecoditioner(self, numVerts: int, numEdges: int, numFaces: int):  # noqa: N802
        self.engine.AllocatePrecoditioner(numVerts, numEdges, numFaces)

    def PreparePreconditioner(self, diagonal, csrOffDiagonals, csrRanges, efSets=None, eeSets=None, vfSets=None,  # noqa: N802
                              efCounts=0, eeCounts=0, vfCounts=0):
        self.engine.PreparePreconditioner(diagonal, csrOffDiagonals, csrRanges, efSets, eeSets, vfSets, efCounts, eeCounts,
                                          vfCounts, phase="begin")
        self._sum_over_ranks(self.engine.exchange_tensor(0))
        self.engine.prepare_end()
        self._apply_exchange = self.engine.exchange_tensor(1)   # sized by this setup's hierarchy

    def Preconditioning(self, z, residual, dim: int = 0):  # noqa: N802
        """z[own vertices] = M^-1 residual; entries of other ranks' vertices are left untouched."""
        if self.p2p:
            return self.engine.Preconditioning(z, residual, dim)
        self.engine.apply_begin(residual)
        self._sum_over_ranks(self._apply_exchange)
        self.engine.apply_end(z)
        return z

    def gather_z(self, z, owned_mask):
        """Debug/test helper: every rank ends with the complete z (sum of the disjoint shards)."""
        import torch
        part = torch.where(owned_mask.unsqueeze(-1), z, torch.zeros_like(z))
        self._sum_over_ranks(part)
        return part

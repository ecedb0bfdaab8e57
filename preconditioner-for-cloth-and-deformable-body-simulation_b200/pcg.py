"""Caller-side PCG over the C ABI (mas_pcg_solve): the loop that calls Preconditioning() once per iteration.

The reference ships no solver — its caller owns the PCG loop (SeSchwarzPreconditioner.h:55-63) — but BASELINE config 2
("1M-vertex cloth, PCG to 1e-5 residual with MAS: iteration count and wall time vs the reference") needs one, and the
north_star parity bar includes "the same PCG iteration counts to convergence".  The whole iteration (block-CSR SpMV,
fused dot/axpy kernels, the MAS apply) runs on the GPU as one replayed CUDA graph; see csrc/mas_pcg.cu.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from .schwarz import SeSchwarzPreconditioner, _is_torch, _ptr


@dataclass
class PcgResult:
    x: object
    iterations: int
    rel_residual: float
    converged: bool
    launches_per_iteration: int


def pcg_solve(pre: SeSchwarzPreconditioner, diagonal, csrOffDiagonals, csrRanges, csrIdx, b, x=None,
              rel_tol: float = 1e-5, max_iter: int = 2000, use_preconditioner: bool = True) -> PcgResult:
    """Solve A x = b (x0 = 0) with MAS-preconditioned CG.  Arrays follow PreparePreconditioner's layout; all of them
    numpy (host) or all of them torch CUDA tensors (device, no copies)."""
    f32 = lambda a: None if _is_torch(a) else np.float32
    i32 = lambda a: None if _is_torch(a) else np.int32
    if x is None:
        if _is_torch(b):
            import torch
            x = torch.zeros_like(b)
        else:
            x = np.zeros_like(np.ascontiguousarray(b, np.float32))
    dp, kind, k0 = _ptr(diagonal, f32(diagonal))
    op, _, k1 = _ptr(csrOffDiagonals, f32(csrOffDiagonals))
    rp, _, k2 = _ptr(csrRanges, i32(csrRanges))
    ip, _, k3 = _ptr(csrIdx, i32(csrIdx))
    bp, kb, k4 = _ptr(b, f32(b))
    xp, kx, k5 = _ptr(x)
    if not (kind == kb == kx):
        raise ValueError("matrix, b and x must live in the same memory space")
    iters, rel = C.c_int(), C.c_float()
    pre._ck(pre.lib.mas_pcg_solve(pre.h, dp, op, rp, ip, bp, xp, C.c_float(rel_tol), int(max_iter), int(bool(use_preconditioner)),
                                  kind, C.byref(iters), C.byref(rel)))
    del k0, k1, k2, k3, k4, k5
    return PcgResult(x, int(iters.value), float(rel.value), bool(pre.get_int(12)), pre.get_int(11))

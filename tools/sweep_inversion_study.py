"""Numerics of the tensor-core inversion that ships (csrc/mas_invert_tc.cuh): block Gauss-Jordan ("sweep") inversion of the
symmetric 96x96 systems by 16-column panels, every panel update one rank-16 GEMM on the tensor cores with 3xTF32 operands.
CPU study, no GPU needed: the algorithm is replayed in numpy with the kernel's operand formats and compared with the parity
bar of tests/test_gpu_parity.py (inverse <= 4x the FP32 oracle's distance from the FP64 inverse + 1e-5, inv r <= 2x + 1e-6)
on the blocks the oracle assembles (fine and Galerkin, k/m = 10 .. 1e5).

Per panel K (columns 16K .. 16K+15) with C = T[:, K], P = T[K, K]^-1 (FP32, un-pivoted Gauss-Jordan on the 16x16 block):
    T[i, j] -= (C P)[i] . C[j]        for i, j outside K        (the tensor-core GEMM, full square)
    T[i, K]  = (C P)[i],  T[K, j] = (C P)[j]^T,  T[K, K] = -P    (same GEMM: operand rows of block K are P and -I, target zeroed)
after six panels T = -A^-1.  The matrix stays symmetric throughout, there is no separate E^T D^-1 E product.

    python tools/sweep_inversion_study.py
"""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
PKG_NAME = "preconditioner-for-cloth-and-deformable-body-simulation_b200"
from tensor_core_tolerance_study import make_product, round_mantissa  # noqa: E402

f32 = np.float32


def invert16(piv):
    """un-pivoted Gauss-Jordan inverse of an SPD 16x16 block in FP32, the order the kernel uses (column p: scale row p,
    eliminate column p from every other row)."""
    n = piv.shape[0]
    M = piv.astype(f32).copy()
    for p in range(n):
        d = f32(1.0) / M[p, p]
        row = (M[p, :] * d).astype(f32)
        row[p] = d
        col = M[:, p].copy()
        for i in range(n):
            if i == p:
                continue
            m = col[i]
            M[i, :] = (M[i, :] - m * row).astype(f32)
            M[i, p] = f32(-m * d)
        M[p, :] = row
    return M


def sweep_invert(A, mode, panel=16, symmetrize_pivot=True):
    prod = make_product(mode)
    n = A.shape[0]
    T = A.astype(f32).copy()
    for i in range(0, n, 3):                       # padding nodes -> identity (cpp:1365-1368)
        if T[i, i] == 0:
            T[i:i + 3, :] = 0
            T[:, i:i + 3] = 0
            T[i:i + 3, i:i + 3] = np.eye(3, dtype=f32)
    for k0 in range(0, n, panel):
        k1 = k0 + panel
        C = T[:, k0:k1].copy()
        piv = C[k0:k1, :].copy()
        if symmetrize_pivot:
            piv = np.tril(piv) + np.tril(piv, -1).T   # the kernel reads the lower triangle of the pivot block
        P = invert16(piv)
        P = ((P + P.T) * f32(0.5)).astype(f32)
        Q = (C @ P).astype(f32)
        Aop = (-Q).astype(f32)
        Bop = C.copy()
        Aop[k0:k1, :] = P
        Bop[k0:k1, :] = -np.eye(panel, dtype=f32)
        T[k0:k1, :] = 0
        T[:, k0:k1] = 0
        T = (T + prod(Aop, Bop.T.copy())).astype(f32)
    inv = (-T).astype(f32)
    return np.tril(inv) + np.tril(inv, -1).T          # the packed layout stores the lower triangle


def main():
    pkg = importlib.import_module(PKG_NAME)
    S = pkg.synth
    from oracle import oracle_binding as ob
    modes = ["fp32", "tf32x3", "tf32"]
    summary = {}
    rng = np.random.RandomState(0)
    cases = [("cloth64 k/m=10", lambda: S.cloth(64, k=10.0)), ("cloth64 k/m=1e3", lambda: S.cloth(64, k=1e3)),
             ("cloth64 k/m=1e5", lambda: S.cloth(64, k=1e5)), ("tet16x16x8", lambda: S.tet_cube(16, 16, 8))]
    print("| mesh | FP32 oracle: inverse / z | " + " | ".join(f"sweep {m}: inverse / z / holds" for m in modes) + " |")
    print("|---|---|" + "---|" * len(modes))
    for name, make in cases:
        mesh = make()
        o32, o64 = ob.OraclePreconditioner("f"), ob.OraclePreconditioner("d")
        for o in (o32, o64):
            o.allocate(mesh)
            o.prepare()
        nb = o32.total_clusters // 32
        nfine = (mesh.nv + 31) // 32
        blocks = sorted(set(list(range(0, nfine, 9)) + list(range(nfine, nb))))
        worst = {m: [0.0, 0.0] for m in modes}
        worst_ref = [0.0, 0.0]
        ok = {m: True for m in modes}
        for b in blocks:
            H = o32.dense_hessian(b).astype(np.float64)
            inv64 = o64.dense_inverse(b).astype(np.float64)
            inv32 = o32.dense_inverse(b).astype(np.float64)
            scale = np.abs(inv64).max()
            e_ref = np.abs(inv32 - inv64).max() / scale
            r = rng.uniform(-1, 1, 96)
            z64 = inv64 @ r
            ez_ref = np.linalg.norm(inv32 @ r - z64) / np.linalg.norm(z64)
            worst_ref = [max(worst_ref[0], e_ref), max(worst_ref[1], ez_ref)]
            for m in modes:
                inv = sweep_invert(H.astype(np.float32), m).astype(np.float64)
                e = np.abs(inv - inv64).max() / scale
                ez = np.linalg.norm(inv @ r - z64) / np.linalg.norm(z64)
                worst[m] = [max(worst[m][0], e), max(worst[m][1], ez)]
                if e > 4 * e_ref + 1e-5 or ez > 2 * ez_ref + 1e-6:
                    ok[m] = False
        summary[name] = {"oracle_fp32": worst_ref, **{m: worst[m] + [ok[m]] for m in modes}}
        cells = " | ".join(f"{worst[m][0]:.1e} / {worst[m][1]:.1e} / {'yes' if ok[m] else 'NO'}" for m in modes)
        print(f"| {name} | {worst_ref[0]:.1e} / {worst_ref[1]:.1e} | {cells} |", flush=True)
    print(json.dumps(summary))


if __name__ == "__main__":
    main()

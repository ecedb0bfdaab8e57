"""Would the batched 96x96 inversion hold the parity tolerance on tensor cores?  (north_star: "tensor cores are used only
if a stated tolerance holds".)  CPU study, no GPU needed.

The inversion (reference cpp:1347-1546: un-pivoted LDL^T by row elimination, inv = E^T D^-1 E) is re-run in numpy on the
dense Hessian blocks the oracle assembles, with the GEMM-shaped work — every rank-1 trailing update and the final
E^T D^-1 E product — fed through an emulated tensor-core input format while pivots, multipliers and accumulation stay
FP32 (what a tcgen05 / mma kernel with FP32 accumulators would do):

  fp32      operands unchanged                      (the CUDA-core kernel that ships)
  tf32      operands rounded to 10 mantissa bits    (kind::tf32)
  tf32x3    a = hi + lo split, hi*hi + hi*lo + lo*hi (error-compensated, 3 MMAs per product)
  bf16x3    same split with 7 mantissa bits          (3 bf16 MMAs)

Bar (tests/test_gpu_parity.py): a dense inverse may be at most 4x as far from the FP64 inverse as the FP32 oracle's is
(+1e-5 of the block's largest entry), and z = inv r at most 2x (+1e-6).  Meshes: the cloth recipe at k/m = 10, 1e3
(BASELINE) and 1e5 (ill-conditioned), fine blocks and coarse (Galerkin) blocks.

    python tools/tensor_core_tolerance_study.py            # prints a markdown table and one JSON line
"""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG_NAME = "preconditioner-for-cloth-and-deformable-body-simulation_b200"


def round_mantissa(x, bits):
    """round-to-nearest-even of float32 values to `bits` explicit mantissa bits (tf32: 10, bf16: 7)."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    drop = 23 - bits
    half = np.uint64(1 << (drop - 1))
    lsb = (u >> np.uint64(drop)) & np.uint64(1)
    u = (u + half - np.uint64(1) + lsb) >> np.uint64(drop) << np.uint64(drop)
    return u.astype(np.uint32).view(np.float32).reshape(np.shape(x))


def make_product(mode):
    """outer/matmul product with FP32 accumulation and the operand format of `mode`."""
    f32 = np.float32
    if mode == "fp32":
        return lambda a, b: (a.astype(f32) @ b.astype(f32)).astype(f32)
    bits = 10 if mode.startswith("tf32") else 7
    if mode in ("tf32", "bf16"):
        return lambda a, b: (round_mantissa(a, bits) @ round_mantissa(b, bits)).astype(f32)

    def split(a, b):
        ah, bh = round_mantissa(a, bits), round_mantissa(b, bits)
        al, bl = round_mantissa(a - ah, bits), round_mantissa(b - bh, bits)
        return ((al @ bh).astype(f32) + (ah @ bl).astype(f32) + (ah @ bh).astype(f32)).astype(f32)
    return split


def invert(A, mode, panel=16):
    """Blocked form of the reference elimination (csrc/mas_assemble.cu): per 16-column panel the diagonal tile is
    eliminated in FP32 exactly like the reference; the panel products and the trailing update go through `mode`."""
    f32 = np.float32
    prod = make_product(mode)
    n = A.shape[0]
    A = A.astype(f32).copy()
    for i in range(0, n, 3):                       # padding nodes -> identity (cpp:1365-1368)
        if A[i, i] == 0:
            A[i:i + 3, :] = 0
            A[:, i:i + 3] = 0
            A[i:i + 3, i:i + 3] = np.eye(3, dtype=f32)
    E = np.eye(n, dtype=f32)                       # accumulates L^-1
    D = np.zeros(n, f32)
    S = A.copy()                                   # Schur complement (lower part is what matters)
    for k0 in range(0, n, panel):
        k1 = min(n, k0 + panel)
        # (a) diagonal tile: FP32 scalar elimination, W = L_kk^-1
        T = S[k0:k1, k0:k1].copy()
        W = np.eye(k1 - k0, dtype=f32)
        for x in range(k1 - k0):
            piv = T[x, x]
            for y in range(x + 1, k1 - k0):
                r = f32(-T[y, x] / piv)
                T[y, :] = (T[y, :] + r * T[x, :]).astype(f32)
                W[y, :] = (W[y, :] + r * W[x, :]).astype(f32)
        D[k0:k1] = np.diag(T)
        # (b) E_K. <- W E_K. ; M = A_.K W^T ; L = M D^-1
        E[k0:k1, :k0] = prod(W, E[k0:k1, :k0])
        E[k0:k1, k0:k1] = W
        if k1 < n:
            M = prod(S[k1:, k0:k1], W.T.copy())
            L = (M / D[k0:k1][None, :]).astype(f32)
            # (c) trailing update and the new column block of E
            S[k1:, k1:] = (S[k1:, k1:] - prod(L, M.T.copy())).astype(f32)
            E[k1:, :k1] = (E[k1:, :k1] - prod(L, E[k0:k1, :k1])).astype(f32)
    Dinv = (f32(1) / D).astype(f32)
    inv = prod(E.T.copy(), (Dinv[:, None] * E).astype(f32))
    return ((inv + inv.T) * f32(0.5)).astype(f32)


def main():
    pkg = importlib.import_module(PKG_NAME)
    S = pkg.synth
    from oracle import oracle_binding as ob
    modes = ["fp32", "tf32", "tf32x3", "bf16x3"]
    rows, summary = [], {}
    rng = np.random.RandomState(0)
    for k in (10.0, 1e3, 1e5):
        mesh = S.cloth(64, k=k)
        o32, o64 = ob.OraclePreconditioner("f"), ob.OraclePreconditioner("d")
        for o in (o32, o64):
            o.allocate(mesh)
            o.prepare()
        nb = o32.total_clusters // 32
        blocks = sorted(set(list(range(0, 128, 9)) + list(range(128, nb))))   # fine sample + every coarse block
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
                inv = invert(H.astype(np.float32), m).astype(np.float64)
                e = np.abs(inv - inv64).max() / scale
                ez = np.linalg.norm(inv @ r - z64) / np.linalg.norm(z64)
                worst[m] = [max(worst[m][0], e), max(worst[m][1], ez)]
                if e > 4 * e_ref + 1e-5 or ez > 2 * ez_ref + 1e-6:
                    ok[m] = False
        cond = float(np.linalg.cond(o32.dense_hessian(0).astype(np.float64)))
        summary[f"k/m={k:g}"] = {"cond_block0": cond, "oracle_fp32": worst_ref, **{m: worst[m] + [ok[m]] for m in modes}}
        rows.append((k, cond, worst_ref, worst, ok))
    print("| k/m | cond(block 0) | FP32 oracle: inverse / z | " + " | ".join(f"{m}: inverse / z / holds" for m in modes) + " |")
    print("|---|---|---|" + "---|" * len(modes))
    for k, cond, wr, w, ok in rows:
        cells = " | ".join(f"{w[m][0]:.1e} / {w[m][1]:.1e} / {'yes' if ok[m] else 'NO'}" for m in modes)
        print(f"| {k:g} | {cond:.1e} | {wr[0]:.1e} / {wr[1]:.1e} | {cells} |")
    print(json.dumps(summary))


if __name__ == "__main__":
    main()

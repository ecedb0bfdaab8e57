"""What would storing the packed inverses in 16 bits cost?  The apply is bound by reading 18,624 B of FP32 inverse per
32-node domain (97 % of its HBM bytes); FP16 or BF16 storage with FP32 arithmetic would halve them.
CPU study, no GPU: the oracle's hierarchy and dense inverses (FP32) are taken as they are, the inverses are rounded to the
storage format, the multilevel apply is re-run in numpy and PCG is iterated to 1e-5 with it.

    python tools/inverse_storage_study.py [n=256]      # n x n cloth; prints a markdown table
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def round_mantissa(x, bits):
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    drop = 23 - bits
    half = np.uint64(1 << (drop - 1))
    lsb = (u >> np.uint64(drop)) & np.uint64(1)
    u = (u + half - np.uint64(1) + lsb) >> np.uint64(drop) << np.uint64(drop)
    return u.astype(np.uint32).view(np.float32).reshape(np.shape(x))


class NumpyApply:
    """A.7 of SURVEY.md in numpy over the oracle's hierarchy: R_l by summation up goingNext, Z_l = blockdiag^-1 R_l,
    z = Z_0 + sum of the prolonged coarse levels (levels 1 .. min(numLevel, 4) - 1)."""

    def __init__(self, o, nv):
        self.nv = nv
        self.s2o = o.sorted_get_original()
        self.tc = o.total_clusters
        self.gn = o.going_next()[:self.tc]
        self.ls = o.level_size()
        self.L = o.num_level
        nb = self.tc // 32
        self.inv = np.stack([o.dense_inverse(b) for b in range(nb)]).astype(np.float32)      # [nb, 96, 96]
        self.ct = o.coarse_tables()

    def with_storage(self, fmt):
        c = object.__new__(NumpyApply)
        c.__dict__.update(self.__dict__)
        if fmt == "fp16":
            c.inv = self.inv.astype(np.float16).astype(np.float32)
        elif fmt == "bf16":
            c.inv = round_mantissa(self.inv, 7)
        elif fmt == "fp16_scaled":
            # one FP32 scale per domain keeps small inverses out of the FP16 subnormal range
            s = np.abs(self.inv).max(axis=(1, 2), keepdims=True)
            s[s == 0] = 1
            c.inv = (self.inv / s).astype(np.float16).astype(np.float32) * s
        return c

    def __call__(self, r4):
        nv, tc = self.nv, self.tc
        R = np.zeros((tc, 3), np.float32)
        R[:nv] = r4[self.s2o, :3]
        for l in range(self.L - 1):
            beg, cnt = (0, nv) if l == 0 else (int(self.ls[l][1]), int(self.ls[l][0]))
            np.add.at(R, self.gn[beg:beg + cnt], R[beg:beg + cnt])
        Z = np.einsum("bij,bj->bi", self.inv, R.reshape(-1, 96)).reshape(tc, 3).astype(np.float32)
        zs = Z[:nv].copy()
        for l in range(1, min(self.L, 4)):
            zs += Z[self.ct[:, l - 1]]
        z = np.zeros((nv, 4), np.float32)
        z[self.s2o, :3] = zs
        return z


def main():
    pkg = importlib.import_module("preconditioner-for-cloth-and-deformable-body-simulation_b200")
    S = pkg.synth
    from oracle import oracle_binding as ob
    from oracle.cpu_pcg import bsr_matrix, cpu_pcg
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    print(f"| mesh | storage | bytes per domain | z rel-L2 vs FP32 storage | PCG iterations to 1e-5 |")
    print("|---|---|---|---|---|")
    for name, mesh in ((f"cloth {n}x{n}, k/m = 1e3", S.cloth(n)), (f"cloth {n}x{n}, k/m = 1e5", S.cloth(n, k=1e5))):
        o = ob.OraclePreconditioner("f")
        o.allocate(mesh)
        o.prepare()
        r = S.residual(mesh.nv)
        base = NumpyApply(o, mesh.nv)
        z_or = o.apply(r)
        z32 = base(r)
        assert np.linalg.norm(z32[:, :3] - z_or[:, :3]) <= 1e-5 * np.linalg.norm(z_or[:, :3])     # the numpy apply is the oracle's
        A = bsr_matrix(mesh)
        for fmt, nbytes in (("fp32", 18624), ("fp16", 9312), ("fp16_scaled", 9316), ("bf16", 9312)):
            ap = base.with_storage(fmt)
            z = ap(r)
            err = np.linalg.norm(z[:, :3] - z32[:, :3]) / np.linalg.norm(z32[:, :3])
            _, its = cpu_pcg(A, r, ap)
            print(f"| {name} | {fmt} | {nbytes:,} | {err:.1e} | {its} |", flush=True)


if __name__ == "__main__":
    main()

"""Full-size value fixtures: the reference's own z at the sizes BASELINE.json quotes (1M-vertex cloth, 1M-vertex tet cube,
2048^2 cloth), sampled, plus the FP64 arbiter's z on the same sample and the reference-preconditioned PCG iteration count.

    python tests/golden/make_golden_full.py [cloth1024 tet128 cloth2048]

Runs the reference's code (oracle/_ref/libmas_ref*.so = SeSchwarzPreconditioner.cpp compiled by oracle/build_ref.sh) and the
plain-C restatement in double precision (oracle/mas_oracle.c, the arbiter of SURVEY 8c) HERE, where /root/reference exists;
the fixtures travel to the GPU box.  Inputs are regenerated from synth.py at test time (seeded): only outputs are stored.
The 2048^2 cloth needs the Q5-fixed reference build (the stock one overruns its buffers beyond 1.08M vertices)."""
import importlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
S = importlib.import_module("preconditioner-for-cloth-and-deformable-body-simulation_b200.synth")
from oracle import oracle_binding as ob  # noqa: E402
from oracle import ref_binding as rb  # noqa: E402
from oracle.cpu_pcg import bsr_matrix, cpu_pcg  # noqa: E402

CASES = {"cloth1024": (2, True), "tet128": (3, True), "cloth2048": (4, False)}   # BASELINE config index, run the CPU PCG too


def main():
    names = sys.argv[1:] or list(CASES)
    threads = len(os.sched_getaffinity(0))
    for name in names:
        cfg, do_pcg = CASES[name]
        t0 = time.time()
        mesh = S.config(cfg)
        big = mesh.nv > 33792 * 32
        p = rb.RefPreconditioner(threads=threads, q5fix=big)
        p.allocate(mesh)
        p.prepare()
        r = S.residual(mesh.nv)
        z_ref = p.apply(r)
        o64 = ob.OraclePreconditioner("d")
        o64.allocate(mesh)
        o64.prepare()
        z64 = o64.apply(r)
        idx = np.arange(0, mesh.nv, 257, dtype=np.int64)
        n64 = np.linalg.norm(z64[:, :3].astype(np.float64))
        out = dict(config=cfg, nv=mesh.nv, idx=idx, z_ref=z_ref[idx], z_f64=z64[idx].astype(np.float64),
                   norm_z_ref=np.linalg.norm(z_ref[:, :3].astype(np.float64)), norm_z_f64=n64,
                   rel_l2_ref_vs_f64=np.linalg.norm((z_ref - z64)[:, :3].astype(np.float64)) / n64,
                   sum_z_ref=z_ref[:, :3].astype(np.float64).sum(0), r_dot_z_ref=float((r[:, :3].astype(np.float64) * z_ref[:, :3]).sum()),
                   level_size=np.asarray(p.level_size()), total_clusters=p.total_clusters, reference_threads=threads,
                   reference_build="q5fix" if big else "stock")
        if do_pcg:
            A = bsr_matrix(mesh)
            _, its = cpu_pcg(A, r, p.apply)
            out["pcg_iterations_reference"] = its
        path = os.path.join(HERE, "full_" + name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "nv", mesh.nv, "levels", p.num_level, "rel_l2(ref, f64)", out["rel_l2_ref_vs_f64"],
              "pcg", out.get("pcg_iterations_reference"), os.path.getsize(path), "bytes", round(time.time() - t0, 1), "s", flush=True)
        del p, o64


if __name__ == "__main__":
    main()

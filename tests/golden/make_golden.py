"""Generates the golden fixtures in this directory by running the reference's own code
(oracle/_ref/libmas_ref.so = SeSchwarzPreconditioner.cpp compiled by oracle/build_ref.sh, one thread).

    python tests/golden/make_golden.py

Needs /root/reference (to have built oracle/_ref); the fixtures travel to the GPU box, the reference does not.
Inputs are regenerated from synth.py at test time (seeded), so only the reference's OUTPUTS are stored."""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
S = importlib.import_module("preconditioner-for-cloth-and-deformable-body-simulation_b200.synth")
from oracle import ref_binding as rb  # noqa: E402

CASES = {
    "cloth64": dict(kind="cloth", n=64),
    "cloth40_collisions": dict(kind="cloth_coll", n=40, ef=100, ee=100, vf=200, seed=7),
    "tet8x8x4": dict(kind="tet", dims=(8, 8, 4)),
}


def build_mesh(spec):
    if spec["kind"] == "cloth":
        return S.cloth(spec["n"])
    if spec["kind"] == "cloth_coll":
        m = S.cloth(spec["n"], with_topology=True)
        return S.add_collisions(m, spec["ef"], spec["ee"], spec["vf"], seed=spec["seed"])
    if spec["kind"] == "tet":
        return S.tet_cube(*spec["dims"])
    raise ValueError(spec)


def main():
    for name, spec in CASES.items():
        mesh = build_mesh(spec)
        p = rb.RefPreconditioner(threads=1)
        p.allocate(mesh)
        p.prepare()
        L, tc = p.num_level, p.total_clusters
        r = S.residual(mesh.nv, 1)
        z = p.apply(r)
        st, mapped = p.stencils()
        nb = tc // 32
        blocks = sorted({0, nb // 2, nb - 1})
        out = dict(
            num_level=L, total_clusters=tc, level_size=p.level_size(),
            morton=p.morton(), sorted_get_original=p.sorted_get_original(),
            going_next=p.going_next(tc), fine_connect_mask=p.fine_connect_mask(),
            coarse_tables=p.coarse_tables()[:, :max(L - 1, 0)],
            coarse_space_tables=np.stack([p.coarse_space_table(l) for l in range(L)]),
            stencils=np.frombuffer(st.tobytes(), np.uint8).reshape(-1, 80), stencil_index_mapped=mapped,
            z=z, inverse_blocks=np.array(blocks, np.int32),
            inverses=np.stack([p.dense_inverse(b) for b in blocks]),
        )
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "levels", L, "clusters", tc, "stencils", len(st), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

"""The proximity stencil producer (collide.py, caller-side workload generation) against brute force, and the
stencils it makes through the oracle: the CPU restatement and the compiled reference build the same 80-byte records and the
same hierarchy from them (geometrically meaningful VF / EE / EF stencils instead of the random ones of synth.add_collisions)."""
import numpy as np
import pytest

from helpers import assert_stencils_equal, make_oracle, rel_l2


@pytest.fixture(scope="module")
def folded(synth, pkg):
    m = synth.folded_cloth(16, 16)
    return pkg.collide.proximity_stencils(m, radius=0.006, device="cpu")


def test_vertex_face_pairs_match_brute_force(folded):
    m = folded
    P, F = m.positions[:, :3].astype(np.float64), m.faces[:, :3]
    want = {}
    for v in range(m.nv):
        for f in range(m.nf):
            if v in F[f]:
                continue
            a, b, c = P[F[f, 0]], P[F[f, 1]], P[F[f, 2]]
            e0, e1, d = a - c, b - c, P[v] - c
            b0, b1 = np.linalg.solve(np.array([[e0 @ e0, e0 @ e1], [e0 @ e1, e1 @ e1]]), np.array([d @ e0, d @ e1]))
            q = c + b0 * e0 + b1 * e1
            dist = np.linalg.norm(P[v] - q)
            if b0 > 0 and b1 > 0 and b0 + b1 < 1 and 1e-7 < dist < 0.006:
                want[(v, f)] = (b0, b1, (P[v] - q) / dist)
    assert m.vf_total == len(want) > 0
    vf = m.vf[m.ef_total + m.ee_total:]
    for rec in vf:
        b0, b1, n = want[(int(rec["vId"]), int(rec["fId"]))]
        assert abs(rec["bary"][0] - b0) < 1e-4 and abs(rec["bary"][1] - b1) < 1e-4 and abs(rec["pad"] - (b0 + b1)) < 1e-4
        assert np.abs(rec["normal"][:3] - n).max() < 1e-3 and rec["normal"][3] == 0


def test_edge_stencils_are_geometrically_consistent(folded):
    m = folded
    P, E, F = m.positions[:, :3].astype(np.float64), m.edges[:, :2], m.faces[:, :3]
    assert m.ee_total > 0 and m.ef_total > 0
    for rec in m.ee[m.ef_total:m.ef_total + m.ee_total]:
        e0, e1 = E[rec["eId0"]], E[rec["eId1"]]
        assert len({*e0.tolist(), *e1.tolist()}) == 4                      # no shared vertex
        c0 = rec["bary"][0] * P[e0[0]] + (1 - rec["bary"][0]) * P[e0[1]]
        c1 = rec["bary"][1] * P[e1[0]] + (1 - rec["bary"][1]) * P[e1[1]]
        d = c0 - c1
        assert 0 < np.linalg.norm(d) < 0.006 + 1e-6
        assert np.abs(d / np.linalg.norm(d) - rec["normal"][:3]).max() < 1e-3
        assert abs(d @ (P[e0[1]] - P[e0[0]])) < 1e-6 and abs(d @ (P[e1[1]] - P[e1[0]])) < 1e-6   # common perpendicular
    for rec in m.ef[:m.ef_total]:
        e, f = E[rec["eId"]], F[rec["fId"]]
        x_edge = rec["bary"][0] * P[e[0]] + (1 - rec["bary"][0]) * P[e[1]]
        x_face = rec["bary"][1] * P[f[0]] + rec["bary"][2] * P[f[1]] + (1 - rec["bary"][1] - rec["bary"][2]) * P[f[2]]
        assert np.abs(x_edge - x_face).max() < 1e-6                        # the intersection point, seen from both sides
        assert 0 < rec["bary"][0] < 1 and min(rec["bary"][1], rec["bary"][2], 1 - rec["bary"][1] - rec["bary"][2]) > 0


def test_oracle_and_reference_agree_on_proximity_stencils(synth, pkg, oracle_lib, ref_lib):
    m = pkg.collide.proximity_stencils(synth.folded_cloth(24, 24), radius=0.006, device="cpu")
    assert m.ef_total + m.ee_total + m.vf_total > 100
    o = make_oracle(oracle_lib, m, "f")
    p = ref_lib.RefPreconditioner(threads=1)
    p.allocate(m)
    p.prepare()
    assert o.stencil_num == p.stencil_num == m.ef_total + m.ee_total + m.vf_total
    so, mo = o.stencils()
    sr, mr = p.stencils()
    assert np.array_equal(mo, mr)
    assert_stencils_equal(so, sr)
    assert np.array_equal(np.asarray(o.level_size()), np.asarray(p.level_size())[:o.num_level + 1])
    assert np.array_equal(o.going_next()[:o.total_clusters], p.going_next(o.total_clusters))
    r = synth.residual(m.nv)
    assert rel_l2(o.apply(r), p.apply(r)) < 1e-4

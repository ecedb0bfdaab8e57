"""Shared comparison helpers for the parity tests."""
import numpy as np


def make_oracle(ob, mesh, precision="f", **kw):
    o = ob.OraclePreconditioner(precision, **kw)
    o.allocate(mesh)
    o.prepare()
    return o


def assert_structure_equal(a, b, nv, check_stencils=True):
    """Bit-exact comparison of every integer structure of two implementations (reference/oracle/GPU wrappers
    expose the same getters).  Covers Morton codes, both permutations, level sizes, goingNext, the per-level
    coarse-space tables, the ancestor table and the level-0 component masks."""
    L = a.num_level
    assert L == b.num_level
    ls = np.asarray(a.level_size())
    assert np.array_equal(ls, np.asarray(b.level_size())), (ls.tolist(), np.asarray(b.level_size()).tolist())
    assert a.total_clusters == b.total_clusters
    assert np.array_equal(a.morton(), b.morton())
    assert np.array_equal(a.sorted_get_original(), b.sorted_get_original())
    assert np.array_equal(a.original_get_sorted(), b.original_get_sorted())
    lo_a, hi_a = a.aabb()
    lo_b, hi_b = b.aabb()
    assert np.array_equal(lo_a, lo_b) and np.array_equal(hi_a, hi_b)
    assert np.array_equal(a.fine_connect_mask(), b.fine_connect_mask())
    tc = a.total_clusters
    ga, gb = np.asarray(a.going_next())[:tc], np.asarray(b.going_next())[:tc]
    for l in range(L):
        beg = 0 if l == 0 else int(ls[l][1])
        cnt = nv if l == 0 else int(ls[l][0])
        assert np.array_equal(ga[beg:beg + cnt], gb[beg:beg + cnt]), f"goingNext differs at level {l}"
        assert np.array_equal(a.coarse_space_table(l), b.coarse_space_table(l)), f"coarse space table {l}"
    assert np.array_equal(a.coarse_tables()[:, :L - 1], b.coarse_tables()[:, :L - 1])
    if check_stencils:
        assert a.stencil_num == b.stencil_num
        if a.stencil_num:
            sa, ma = a.stencils()
            sb, mb = b.stencils()
            assert np.array_equal(ma, mb)
            assert_stencils_equal(sa, sb)


def assert_stencils_equal(sa, sb):
    a = np.frombuffer(sa.tobytes(), np.uint8).reshape(-1, 80)
    b = np.frombuffer(sb.tobytes(), np.uint8).reshape(-1, 80)
    for lo, hi in ((0, 8), (48, 52), (64, 80)):   # n, nFirst | stiff | direction
        assert np.array_equal(a[:, lo:hi], b[:, lo:hi])
    n = a[:, 0]
    for k in range(5):                             # index[k], weight[k] are defined for k < n only
        m = n > k
        assert np.array_equal(a[m, 8 + 4 * k:12 + 4 * k], b[m, 8 + 4 * k:12 + 4 * k])
        assert np.array_equal(a[m, 28 + 4 * k:32 + 4 * k], b[m, 28 + 4 * k:32 + 4 * k])


def rel_l2(a, b):
    a = np.asarray(a, np.float64)[..., :3]
    b = np.asarray(b, np.float64)[..., :3]
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def arbiter_ok(z_test, z_ref32, z_f64, slack=2.0, floor=1e-6):
    """SURVEY §8c tolerance: the implementation under test may be at most `slack` times as far from the FP64
    arbiter as the reference-arithmetic (FP32) result is, plus a 1e-6 relative floor."""
    nz = np.linalg.norm(np.asarray(z_f64, np.float64)[..., :3])
    e_test = rel_l2(z_test, z_f64) * nz
    e_ref = rel_l2(z_ref32, z_f64) * nz
    return e_test <= slack * e_ref + floor * nz, e_test / nz, e_ref / nz


from oracle.cpu_pcg import bsr_matrix, cpu_pcg  # noqa: E402,F401  (test-infrastructure PCG loop)

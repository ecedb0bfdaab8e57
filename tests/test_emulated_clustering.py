"""The clustering kernels of csrc/mas_cluster.cu — connect_mask_l0 / _lx, collision_connect, close_components, exclusive_scan,
number_components, next_level_table, coarse_tables — run WITHOUT a GPU: tests/emu/cluster_emu.cpp includes the .cu file
itself (host launches guarded out) and plays every thread block with OS threads (tests/emu/cuda_emu.h), in the launch order
of build_hierarchy.  Integer work: level sizes, goingNext, the level-0 component masks, every coarse-space table and the
ancestor table must equal the oracle's bit for bit (the oracle is pinned to the compiled reference by
tests/test_oracle_vs_reference.py); under ThreadSanitizer the atomicOr / scan / flood-fill code must be race-free."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from helpers import make_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"


def _build(tmp, extra=()):
    if not shutil.which("g++") or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    exe = str(tmp / "cluster_emu")
    p = subprocess.run(["g++", "-std=c++20", "-O1", "-g", "-pthread", "-ffp-contract=off", "-w", *extra, "-I", CUDA_INC, "-I",
                        os.path.join(ROOT, "tests", "emu"), os.path.join(ROOT, "tests", "emu", "cluster_emu.cpp"), "-o", exe],
                       capture_output=True, text=True)
    return exe, p


def _run(exe, o, mesh, env=None):
    starts, idx = o.sorted_adjacency()
    parts = [np.array([mesh.nv, mesh.nnz, o.stencil_num], np.int32), starts.astype(np.int32), idx.astype(np.int32)]
    if o.stencil_num:
        rec, mapped = o.stencils()
        parts += [np.frombuffer(rec.tobytes(), np.uint8), np.ascontiguousarray(mapped, np.int32)]
    p = subprocess.run([exe], input=b"".join(x.tobytes() for x in parts), capture_output=True, timeout=1800, env=env)
    assert p.returncode == 0, p.stderr[-800:]
    a, nv = np.frombuffer(p.stdout, np.int32), mesh.nv
    L, total, k = int(a[0]), int(a[1]), 2
    out = {"L": L, "total": total}
    out["level_size"] = a[k:k + 2 * (L + 1)].reshape(L + 1, 2); k += 2 * (L + 1)
    out["going_next"] = a[k:k + total]; k += total
    out["fine_mask"] = a[k:k + nv].view(np.uint32); k += nv
    out["cst"] = a[k:k + L * nv].reshape(L, nv); k += L * nv
    out["coarse_tables"] = a[k:k + 4 * nv].reshape(nv, 4)
    return out, p.stderr.decode(errors="replace")


@pytest.fixture(scope="module")
def emulator(tmp_path_factory):
    exe, p = _build(tmp_path_factory.mktemp("emu_cluster"))
    assert p.returncode == 0, p.stderr[-2000:]
    return exe


def _collisions(s, n, seed):
    m = s.cloth(n, with_topology=True)
    return s.add_collisions(m, m.nv // 8, m.nv // 8, m.nv // 4, seed=seed)


CASES = {
    "cloth50_ragged": lambda s: s.cloth(50),
    "cloth48_collisions": lambda s: _collisions(s, 48, 7),
    "cloud1500_irregular": lambda s: s.random_cloud(1500, 5, 3),
    "dust1025_nothing_aggregates": lambda s: s.dust(1025),
    "chain100_fragmented_two_levels": lambda s: s.chain(100),
    "rippled48_fragmented": lambda s: s.rippled_cloth(48, amplitude=1e-3),
}


@pytest.mark.parametrize("name", list(CASES))
def test_emulated_clustering_is_bit_exact(name, emulator, synth, oracle_lib):
    mesh = CASES[name](synth)
    o = make_oracle(oracle_lib, mesh)
    got, _ = _run(emulator, o, mesh)
    L = o.num_level
    assert got["L"] == L and got["total"] == o.total_clusters
    ls = np.asarray(o.level_size())
    assert np.array_equal(got["level_size"], ls)
    assert np.array_equal(got["fine_mask"], o.fine_connect_mask())
    want = o.going_next()[:o.total_clusters]
    for l in range(L):                                   # padding slots of goingNext are unspecified: compare live nodes
        beg, cnt = (0, mesh.nv) if l == 0 else (int(ls[l][1]), int(ls[l][0]))
        assert np.array_equal(got["going_next"][beg:beg + cnt], want[beg:beg + cnt]), f"goingNext level {l}"
        assert np.array_equal(got["cst"][l], o.coarse_space_table(l)), f"coarse space table {l}"
    assert np.array_equal(got["coarse_tables"][:, :L - 1], o.coarse_tables()[:, :L - 1])


def test_emulated_clustering_has_no_race_under_thread_sanitizer(tmp_path, synth, oracle_lib):
    exe, p = _build(tmp_path, extra=("-fsanitize=thread",))
    if p.returncode != 0:
        pytest.skip("ThreadSanitizer runtime not available: " + p.stderr[-200:])
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=0")
    for mesh in (synth.chain(100), _collisions(synth, 24, 3)):
        _, err = _run(exe, make_oracle(oracle_lib, mesh), mesh, env)
        assert err.count("WARNING: ThreadSanitizer: data race") == 0, err[:2000]

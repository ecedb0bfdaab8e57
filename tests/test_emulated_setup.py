"""The setup kernels of csrc/mas_assemble.cu — cross_bank, fine_assemble_invert (gather + blocked inversion), carry_up and
coarse_invert — run WITHOUT a GPU: tests/emu/assemble_emu.cpp includes the .cu file itself (host launches guarded out) and
plays every thread block with OS threads, atomics included (tests/emu/cuda_emu.h).  Inputs are the caller's Hessian and the
oracle's hierarchy; every resulting inverse, fine and Galerkin, is held to the GPU parity test's bar against the FP64 oracle.
Under ThreadSanitizer the same run checks the gather's "one writer per tile element" design and every barrier of the setup
path for data races."""
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
    exe = str(tmp / "assemble_emu")
    p = subprocess.run(["g++", "-std=c++20", "-O1", "-g", "-pthread", "-ffp-contract=off", "-w", *extra, "-I", CUDA_INC, "-I",
                        os.path.join(ROOT, "tests", "emu"), os.path.join(ROOT, "tests", "emu", "assemble_emu.cpp"), "-o", exe],
                       capture_output=True, text=True)
    return exe, p


def _run(exe, o, mesh, env=None):
    tc = o.total_clusters
    starts, idx = o.sorted_adjacency()
    parts = [np.array([mesh.nv, o.num_level, tc, mesh.nnz], np.int32), np.ascontiguousarray(o.level_size(), np.int32),
             o.sorted_get_original().astype(np.int32), starts.astype(np.int32), idx.astype(np.int32), o.going_next()[:tc].astype(np.int32),
             np.asarray(mesh.nbr_starts, np.int32), np.ascontiguousarray(mesh.diag, np.float32), np.ascontiguousarray(mesh.offdiag, np.float32),
             np.array([o.stencil_num], np.int32)]
    if o.stencil_num:
        st, mapped = o.stencils()
        parts += [np.frombuffer(st.tobytes(), np.uint8), np.ascontiguousarray(mapped, np.int32)]
    p = subprocess.run([exe], input=b"".join(x.tobytes() for x in parts), capture_output=True, timeout=1800, env=env)
    assert p.returncode == 0, p.stderr[-800:]
    return np.frombuffer(p.stdout, np.float32).reshape(tc // 32, 96, 96), p.stderr.decode(errors="replace")


@pytest.fixture(scope="module")
def emulator(tmp_path_factory):
    exe, p = _build(tmp_path_factory.mktemp("emu_setup"))
    assert p.returncode == 0, p.stderr[-2000:]
    return exe


CASES = {
    "tet8x8x4": lambda s: s.tet_cube(8, 8, 4),
    "cloth24_duplicate_edges": lambda s: s.cloth_with_duplicate_edges(24),
    "cloud900_irregular": lambda s: s.random_cloud(900, 3, 5),
    "cloth40_skewed_blocks": lambda s: s.cloth(40, skew=0.05),
    "chain100_fragmented": lambda s: s.chain(100),
    "cloth40_collisions": lambda s: s.add_collisions(s.cloth(40, with_topology=True), 100, 100, 200, seed=7),
}


@pytest.mark.parametrize("name", list(CASES))
def test_emulated_setup_inverts_every_block_within_the_bar(name, emulator, synth, oracle_lib):
    mesh = CASES[name](synth)
    o32, o64 = make_oracle(oracle_lib, mesh, "f"), make_oracle(oracle_lib, mesh, "d")
    inv, _ = _run(emulator, o32, mesh)
    nb = o32.total_clusters // 32
    i64 = np.stack([o64.dense_inverse(b) for b in range(nb)])
    i32 = np.stack([o32.dense_inverse(b) for b in range(nb)])
    scale = np.abs(i64).max(axis=(1, 2))
    e_emu = np.abs(inv - i64).max(axis=(1, 2)) / scale
    e_ref = np.abs(i32 - i64).max(axis=(1, 2)) / scale
    assert np.all(e_emu <= 4 * e_ref + 1e-5), (int(np.argmax(e_emu / (4 * e_ref + 1e-5))), e_emu.max(), e_ref.max())
    assert np.array_equal(inv, inv.transpose(0, 2, 1))


def test_emulated_setup_has_no_race_under_thread_sanitizer(tmp_path, synth, oracle_lib):
    exe, p = _build(tmp_path, extra=("-fsanitize=thread",))
    if p.returncode != 0:
        pytest.skip("ThreadSanitizer runtime not available: " + p.stderr[-200:])
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=0")
    for mesh in (synth.tet_cube(8, 8, 4), synth.cloth_with_duplicate_edges(24),
                 synth.add_collisions(synth.cloth(24, with_topology=True), 60, 60, 120, seed=3)):
        _, err = _run(exe, make_oracle(oracle_lib, mesh, "f"), mesh, env)
        assert err.count("WARNING: ThreadSanitizer: data race") == 0, err[:2000]

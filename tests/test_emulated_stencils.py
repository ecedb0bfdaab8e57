"""The collision-stencil kernels of csrc/mas_cluster.cu (stencil_flag, exclusive_scan, stencil_build: EF -> 5-vertex, EE and
VF -> 4-vertex stencils, signed barycentric weights, sorted-space indices; the reference's literal Q2/Q3 indexing) run
WITHOUT a GPU (tests/emu/stencil_emu.cpp includes the .cu file itself) and must reproduce the oracle's stencil records and
mapped indices byte for byte, in input order."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from helpers import assert_stencils_equal, make_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"


@pytest.fixture(scope="module")
def emulator(tmp_path_factory):
    if not shutil.which("g++") or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    exe = str(tmp_path_factory.mktemp("emu_stencil") / "stencil_emu")
    subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-ffp-contract=off", "-w", "-I", CUDA_INC, "-I", os.path.join(ROOT, "tests", "emu"),
                    os.path.join(ROOT, "tests", "emu", "stencil_emu.cpp"), "-o", exe], check=True)
    return exe


@pytest.mark.parametrize("n,n_ef,n_ee,n_vf,seed", [(40, 100, 100, 200, 7), (33, 2000, 2000, 2000, 22), (24, 0, 50, 0, 3), (24, 40, 0, 90, 4)])
def test_emulated_stencil_build_matches_the_oracle(n, n_ef, n_ee, n_vf, seed, emulator, synth, oracle_lib):
    mesh = synth.add_collisions(synth.cloth(n, with_topology=True), n_ef, n_ee, n_vf, seed=seed)
    o = make_oracle(oracle_lib, mesh)
    raw = lambda a: np.frombuffer(np.ascontiguousarray(a).tobytes(), np.uint8)
    hdr = np.array([mesh.nv, mesh.ne, mesh.nf, mesh.ef_total, mesh.ee_total, mesh.vf_total, 0, mesh.ef.shape[0], mesh.ee.shape[0],
                    mesh.vf.shape[0]], np.int32)
    parts = [hdr, np.ascontiguousarray(mesh.edges, np.int32), np.ascontiguousarray(mesh.faces, np.int32),
             o.original_get_sorted().astype(np.int32), raw(mesh.ef), raw(mesh.ee), raw(mesh.vf)]
    p = subprocess.run([emulator], input=b"".join(x.tobytes() for x in parts), capture_output=True, timeout=900, check=True)
    count = int(np.frombuffer(p.stdout, np.int32, 1)[0])
    assert count == o.stencil_num > 0
    rec = np.frombuffer(p.stdout, np.dtype((np.void, 80)), count, 4)
    mapped = np.frombuffer(p.stdout, np.int32, 5 * count, 4 + 80 * count).reshape(count, 5)
    want_rec, want_mapped = o.stencils()
    assert np.array_equal(mapped, want_mapped)
    assert_stencils_equal(rec, want_rec)

"""The kernels of the device PCG harness (csrc/mas_pcg.cu: sliced-ELL conversion of the caller's block CSR, spmv_dot, axpy_rr,
dot, update_p with the device-side stopping test) run WITHOUT a GPU (tests/emu/pcg_emu.cpp includes the .cu file itself) as
plain CG for a fixed number of iterations, against the same loop in numpy / scipy (oracle/cpu_pcg.py conventions: FP32
vectors, FP64 dot products, x0 = 0)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from helpers import bsr_matrix

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"


@pytest.fixture(scope="module")
def emulator(tmp_path_factory):
    if not shutil.which("g++") or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    exe = str(tmp_path_factory.mktemp("emu_pcg") / "pcg_emu")
    subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-ffp-contract=off", "-w", "-I", CUDA_INC, "-I", os.path.join(ROOT, "tests", "emu"),
                    os.path.join(ROOT, "tests", "emu", "pcg_emu.cpp"), "-o", exe], check=True)
    return exe


def _cg(A, b, iters, tol):
    """plain CG, the harness's arithmetic: FP32 vectors and updates, FP64 dot products"""
    nv = b.shape[0]
    dot = lambda u, v: float(np.dot(u.astype(np.float64), v.astype(np.float64)))
    x = np.zeros(3 * nv, np.float32)
    r = np.ascontiguousarray(b[:, :3], np.float32).reshape(-1).copy()
    p = r.copy()
    rz = rr0 = dot(r, r)
    done, it, Ap = False, 0, None
    for _ in range(iters):
        if done:
            break
        Ap = (A @ p).astype(np.float32)
        alpha = np.float32(rz / dot(p, Ap))
        x += alpha * p
        r -= alpha * Ap
        rr = dot(r, r)
        p = r + np.float32(rr / rz) * p
        rz = rr
        it += 1
        done = rr < tol * tol * rr0
    return x.reshape(nv, 3), r.reshape(nv, 3), Ap.reshape(nv, 3), it, done


@pytest.mark.parametrize("name,iters", [("cloth20", 6), ("tet8x8x4_ragged_rows", 5), ("cloud700_irregular", 4), ("cloth12_converges", 40)])
def test_emulated_cg_matches_numpy(name, iters, emulator, synth):
    mesh = {"cloth20": lambda: synth.cloth(20), "tet8x8x4_ragged_rows": lambda: synth.tet_cube(8, 8, 4),
            "cloud700_irregular": lambda: synth.random_cloud(700, 8, 12), "cloth12_converges": lambda: synth.cloth(12, k=1.0)}[name]()
    b = synth.residual(mesh.nv, 4)
    tol = 1e-5
    parts = [np.array([mesh.nv, mesh.nnz, iters], np.int32), np.array([tol], np.float32), np.asarray(mesh.nbr_starts, np.int32),
             np.asarray(mesh.nbr_idx, np.int32), np.ascontiguousarray(mesh.diag, np.float32), np.ascontiguousarray(mesh.offdiag, np.float32),
             np.ascontiguousarray(b, np.float32)]
    p = subprocess.run([emulator], input=b"".join(x.tobytes() for x in parts), capture_output=True, timeout=1800, check=True)
    nv = mesh.nv
    x = np.frombuffer(p.stdout, np.float32, 4 * nv, 0).reshape(nv, 4)
    r = np.frombuffer(p.stdout, np.float32, 4 * nv, 16 * nv).reshape(nv, 4)
    Ap = np.frombuffer(p.stdout, np.float32, 4 * nv, 32 * nv).reshape(nv, 4)
    rr, rr0 = np.frombuffer(p.stdout, np.float64, 2, 48 * nv)
    its, done = np.frombuffer(p.stdout, np.int32, 2, 48 * nv + 16)
    xr, rres, Apr, it_ref, done_ref = _cg(bsr_matrix(mesh), b, iters, tol)
    rel = lambda a, c: float(np.linalg.norm(a - c) / max(np.linalg.norm(c), 1e-30))
    if name == "cloth12_converges":
        assert done == 1 and done_ref and abs(int(its) - it_ref) <= 1 and its < iters      # the device-side stopping test fires
        assert rel(x[:, :3], xr) < 1e-3
    else:
        assert its == iters == it_ref and done == 2                                            # stopped by the iteration limit
        assert rel(Ap[:, :3], Apr) < 1e-5 and rel(x[:, :3], xr) < 1e-5 and rel(r[:, :3], rres) < 1e-4
        assert abs(rr0 - float(np.dot(b[:, :3].astype(np.float64).ravel(), b[:, :3].astype(np.float64).ravel()))) <= 1e-12 * rr0
    assert np.all(x[:, 3] == 0) and np.all(Ap[:, 3] == 0)

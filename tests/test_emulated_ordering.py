"""The ordering kernels of csrc/mas_order.cu — the AABB reduction, the 63-bit Morton codes (IEEE sub / div / mul, comparison
clamp, truncating conversion), the inverse permutation and the adjacency in sorted space — run WITHOUT a GPU
(tests/emu/order_emu.cpp includes the .cu file itself; the CUB sort and scan are std::stable_sort and a prefix sum).  Integer
work, bit-exact against the oracle, which tests/test_oracle_vs_reference.py pins to the compiled reference."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from helpers import make_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"


@pytest.fixture(scope="module")
def emulator(tmp_path_factory):
    if not shutil.which("g++") or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    exe = str(tmp_path_factory.mktemp("emu_order") / "order_emu")
    subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-ffp-contract=off", "-w", "-I", CUDA_INC, "-I", os.path.join(ROOT, "tests", "emu"),
                    os.path.join(ROOT, "tests", "emu", "order_emu.cpp"), "-o", exe], check=True)
    return exe


CASES = {
    "cloth64": lambda s: s.cloth(64),
    "cloth_rect96x40": lambda s: s.cloth_rect(96, 40),                     # per-axis normalisation on a non-square sheet
    "tet16x16x8": lambda s: s.tet_cube(16, 16, 8),
    "cloud4000_irregular": lambda s: s.random_cloud(4000, 7, 4),
    "stacked3x20_equal_codes": lambda s: s.stacked_cloth(20, 3),           # ties: ascending original index
    "rippled64": lambda s: s.rippled_cloth(64),
    "chain1_single_vertex": lambda s: s.chain(1),
}


@pytest.mark.parametrize("name", list(CASES))
def test_emulated_ordering_is_bit_exact(name, emulator, synth, oracle_lib):
    mesh = CASES[name](synth)
    o = make_oracle(oracle_lib, mesh)
    nv, nnz = mesh.nv, mesh.nnz
    parts = [np.array([nv, nnz], np.int32), np.ascontiguousarray(mesh.positions, np.float32), np.asarray(mesh.nbr_starts, np.int32),
             np.asarray(mesh.nbr_idx, np.int32)]
    p = subprocess.run([emulator], input=b"".join(x.tobytes() for x in parts), capture_output=True, timeout=900, check=True)
    buf, k = p.stdout, 0
    aabb = np.frombuffer(buf, np.float32, 8, k); k += 32
    code = np.frombuffer(buf, np.uint64, nv, k); k += 8 * nv
    s2o = np.frombuffer(buf, np.int32, nv, k); k += 4 * nv
    o2s = np.frombuffer(buf, np.int32, nv, k); k += 4 * nv
    adj_start = np.frombuffer(buf, np.int32, nv + 1, k); k += 4 * (nv + 1)
    adj_idx = np.frombuffer(buf, np.int32, nnz, k)
    lo, hi = o.aabb()
    assert np.array_equal(aabb[:4], lo) and np.array_equal(aabb[4:], hi)
    assert np.array_equal(code, o.morton())
    assert np.array_equal(s2o, o.sorted_get_original()) and np.array_equal(o2s, o.original_get_sorted())
    want_start, want_idx = o.sorted_adjacency()
    assert np.array_equal(adj_start, want_start) and np.array_equal(adj_idx, want_idx)

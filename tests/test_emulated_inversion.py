"""The batched 96x96 inversion kernels, run WITHOUT a GPU: tests/emu/invert_emu.cpp includes csrc/mas_invert.cuh — the text
the CUDA kernels compile — and plays one thread block with OS threads (barriers, warp shuffles and the m16n8k8 TF32 MMA
fragment exchange emulated in tests/emu/cuda_emu.h; fragment layouts as in the PTX ISA / CUTLASS SM80_16x8x8_F32TF32TF32F32_TN).
Every MAS_OPT_INVERT_VARIANT is inverted against the FP64 inverse of the oracle's assembled blocks with the bar of the GPU
parity test (4x the FP32 oracle's own distance + 1e-5).  This checks logic and indexing of the device code — including the
experimental variants written while no GPU was available — not synchronisation (data races are invisible here) and not speed."""
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
    exe = str(tmp_path_factory.mktemp("emu") / "invert_emu")
    subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-ffp-contract=off", "-I", CUDA_INC, "-I", os.path.join(ROOT, "tests", "emu"),
                    os.path.join(ROOT, "tests", "emu", "invert_emu.cpp"), "-o", exe], check=True)

    def run(variant, dense):
        dense = np.ascontiguousarray(dense, np.float32)
        p = subprocess.run([exe, str(variant)], input=np.int32(dense.shape[0]).tobytes() + dense.tobytes(), capture_output=True,
                           timeout=600, check=True)
        return np.frombuffer(p.stdout, np.float32).reshape(dense.shape)
    return run


@pytest.fixture(scope="module")
def blocks(synth, oracle_lib):
    """Assembled 96x96 systems of a stiff 64^2 cloth with collisions: two fine blocks, a fine block with padding nodes
    (identity rows), level-1 and level-2 Galerkin blocks."""
    m = synth.cloth(50, k=1e5, with_topology=True)                     # 2,500 vertices: the last fine bank is padded
    mesh = synth.add_collisions(m, 150, 150, 300)
    o32, o64 = make_oracle(oracle_lib, mesh, "f"), make_oracle(oracle_lib, mesh, "d")
    nb = o32.total_clusters // 32
    ids = [0, 41, 78, 79, 80, nb - 2, nb - 1]
    H = np.stack([o32.dense_hessian(b) for b in ids]).astype(np.float32)
    return H, np.stack([o64.dense_inverse(b) for b in ids]), np.stack([o32.dense_inverse(b) for b in ids])


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4])
def test_emulated_kernel_inverts_within_the_parity_bar(variant, emulator, blocks):
    H, inv64, inv32 = blocks
    out = emulator(variant, H)
    assert np.array_equal(out, out.transpose(0, 2, 1))                # one stored value per symmetric pair
    scale = np.abs(inv64).max(axis=(1, 2))
    e_gpu = np.abs(out - inv64).max(axis=(1, 2)) / scale
    e_ref = np.abs(inv32 - inv64).max(axis=(1, 2)) / scale
    assert np.all(e_gpu <= 4 * e_ref + 1e-5), (variant, e_gpu, e_ref)


def test_register_factorisation_is_bit_identical_to_the_shipped_kernel(emulator, blocks):
    H = blocks[0]
    assert np.array_equal(emulator(0, H), emulator(1, H))             # bit 0 changes where the arithmetic happens, not what
    assert np.array_equal(emulator(2, H), emulator(3, H))


def test_no_shared_memory_race_under_thread_sanitizer(tmp_path):
    """The same emulation built with -fsanitize=thread: CUDA threads are OS threads and __syncthreads / __syncwarp / named
    barriers are the only ordering between them, exactly the CUDA memory model for shared memory, so a missing barrier in the
    device code shows up as a ThreadSanitizer data race (checked by deleting one: 110 reports).  All variants must be clean."""
    if not shutil.which("g++") or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    exe = str(tmp_path / "invert_emu_tsan")
    build = subprocess.run(["g++", "-std=c++20", "-O1", "-g", "-pthread", "-fsanitize=thread", "-ffp-contract=off", "-I", CUDA_INC, "-I",
                            os.path.join(ROOT, "tests", "emu"), os.path.join(ROOT, "tests", "emu", "invert_emu.cpp"), "-o", exe],
                           capture_output=True, text=True)
    if build.returncode != 0:
        pytest.skip("ThreadSanitizer runtime not available: " + build.stderr[-200:])
    rng = np.random.RandomState(0)
    b = rng.randn(96, 96)
    a = (b @ b.T + 96 * np.eye(96)).astype(np.float32)
    a[93:, :] = 0
    a[:, 93:] = 0                                                       # one padding node (identity path)
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=0")
    for variant in range(5):
        p = subprocess.run([exe, str(variant)], input=np.int32(1).tobytes() + a.tobytes(), capture_output=True, timeout=900, env=env)
        assert p.returncode == 0, p.stderr[-500:]
        races = p.stderr.decode(errors="replace").count("WARNING: ThreadSanitizer: data race")
        assert races == 0, (variant, races, p.stderr.decode(errors="replace")[:1500])

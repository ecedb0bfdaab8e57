"""The batched 96x96 inversion, run WITHOUT a GPU.

  * the FP32 CUDA-core kernel (MAS_OPT_INVERT_VARIANT = 1): tests/emu/invert_emu.cpp includes csrc/mas_invert.cuh — the text
    the CUDA kernel compiles — and plays one thread block with OS threads (barriers and warp shuffles emulated in
    tests/emu/cuda_emu.h);
  * the default tensor-core kernel (csrc/mas_invert_tc.cuh) needs tcgen05 hardware; its ALGORITHM (block Gauss-Jordan by
    16-column panels, 3xTF32 operands, FP32 pivot inverses) is replayed in numpy by tools/sweep_inversion_study.py and held
    to the same bar here.
Bar = the GPU parity test's: at most 4x the FP32 oracle's own distance from the FP64 inverse + 1e-5."""
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

    def run(dense):
        dense = np.ascontiguousarray(dense, np.float32)
        p = subprocess.run([exe], input=np.int32(dense.shape[0]).tobytes() + dense.tobytes(), capture_output=True,
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


def test_emulated_kernel_inverts_within_the_parity_bar(emulator, blocks):
    H, inv64, inv32 = blocks
    out = emulator(H)
    assert np.array_equal(out, out.transpose(0, 2, 1))                # one stored value per symmetric pair
    scale = np.abs(inv64).max(axis=(1, 2))
    e_gpu = np.abs(out - inv64).max(axis=(1, 2)) / scale
    e_ref = np.abs(inv32 - inv64).max(axis=(1, 2)) / scale
    assert np.all(e_gpu <= 4 * e_ref + 1e-5), (e_gpu, e_ref)


@pytest.mark.parametrize("mode", ["fp32", "tf32x3"])
def test_sweep_algorithm_of_the_tensor_core_kernel_holds_the_parity_bar(mode, blocks):
    """numpy replay of csrc/mas_invert_tc.cuh (tools/sweep_inversion_study.py): stiff cloth with collisions, fine blocks, a
    padded block and Galerkin blocks.  Plain TF32 operands would miss the bar by three orders of magnitude (checked there)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from sweep_inversion_study import sweep_invert
    H, inv64, inv32 = blocks
    out = np.stack([sweep_invert(h, mode) for h in H])
    scale = np.abs(inv64).max(axis=(1, 2))
    e_gpu = np.abs(out - inv64).max(axis=(1, 2)) / scale
    e_ref = np.abs(inv32 - inv64).max(axis=(1, 2)) / scale
    assert np.all(e_gpu <= 4 * e_ref + 1e-5), (e_gpu, e_ref)
    bad = np.stack([sweep_invert(h, "tf32") for h in H[:2]])
    assert np.any(np.abs(bad - inv64[:2]).max(axis=(1, 2)) / scale[:2] > 4 * e_ref[:2] + 1e-5)


def test_no_shared_memory_race_under_thread_sanitizer(tmp_path):
    """The same emulation built with -fsanitize=thread: CUDA threads are OS threads and __syncthreads / __syncwarp / named
    barriers are the only ordering between them, exactly the CUDA memory model for shared memory, so a missing barrier in the
    device code shows up as a ThreadSanitizer data race (checked by deleting one: 110 reports)."""
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
    p = subprocess.run([exe], input=np.int32(1).tobytes() + a.tobytes(), capture_output=True, timeout=900, env=env)
    assert p.returncode == 0, p.stderr[-500:]
    races = p.stderr.decode(errors="replace").count("WARNING: ThreadSanitizer: data race")
    assert races == 0, (races, p.stderr.decode(errors="replace")[:1500])

"""The apply kernels of csrc/mas_apply.cu — restrict_fine, restrict_l1, restrict_top, solve_coarse, prolong_sum and the
HBM-bound solve_fine — run WITHOUT a GPU: tests/emu/apply_emu.cpp includes the .cu file itself (host launches guarded out)
and plays every thread block with OS threads (tests/emu/cuda_emu.h).  Inputs are the oracle's hierarchy and its dense
inverses, packed by the same packed_pos() the library uses; the emulated z must satisfy the bar of the GPU parity test
against the oracle's own apply.  Checks the device code's logic (lane-slot layout, shuffled symmetric halves, group sums,
level walk), not synchronisation timing and not speed."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from helpers import arbiter_ok, make_oracle, rel_l2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"


def _build(tmp, extra=()):
    if not shutil.which("g++") or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    exe = str(tmp / "apply_emu")
    p = subprocess.run(["g++", "-std=c++20", "-O1", "-g", "-pthread", "-ffp-contract=off", "-w", *extra, "-I", CUDA_INC, "-I",
                        os.path.join(ROOT, "tests", "emu"), os.path.join(ROOT, "tests", "emu", "apply_emu.cpp"), "-o", exe],
                       capture_output=True, text=True)
    return exe, p


def _input(o, mesh, r):
    tc = o.total_clusters
    parts = [np.array([mesh.nv, o.num_level, tc], np.int32), np.ascontiguousarray(o.level_size(), np.int32),
             o.sorted_get_original().astype(np.int32), o.going_next()[:tc].astype(np.int32)]
    parts += [o.dense_inverse(b).astype(np.float32) for b in range(tc // 32)]
    parts.append(np.ascontiguousarray(r, np.float32))
    parts.append(np.ascontiguousarray(o.coarse_tables(), np.int32))
    return b"".join(p.tobytes() for p in parts)


@pytest.fixture(scope="module")
def emulator(tmp_path_factory):
    exe, p = _build(tmp_path_factory.mktemp("emu_apply"))
    assert p.returncode == 0, p.stderr[-2000:]
    return exe


CASES = {
    "cloth5_one_level": lambda s: s.cloth(5),
    "cloth64_three_levels": lambda s: s.cloth(64),
    "cloth50_ragged": lambda s: s.cloth(50),
    "cloud900_two_levels_multi_bank_top": lambda s: s.random_cloud(900, 3, 5),
    "rippled64_fragmented": lambda s: s.rippled_cloth(64),
    "stacked2x24_ties": lambda s: s.stacked_cloth(24, 2),
    "dust1025_no_edges": lambda s: s.dust(1025),
    "cloth182_four_levels": lambda s: s.cloth(182),                      # 33,124 vertices: restrict_top runs
}


@pytest.mark.parametrize("name", list(CASES))
def test_emulated_apply_matches_the_oracle(name, emulator, synth, oracle_lib):
    mesh = CASES[name](synth)
    o32, o64 = make_oracle(oracle_lib, mesh, "f"), make_oracle(oracle_lib, mesh, "d")
    r = synth.residual(mesh.nv, 2)
    p = subprocess.run([emulator], input=_input(o32, mesh, r), capture_output=True, timeout=900, check=True)
    z = np.frombuffer(p.stdout, np.float32).reshape(mesh.nv, 4)
    z32, z64 = o32.apply(r), o64.apply(r)
    assert np.all(z[:, 3] == 0)
    ok, e_emu, e_ref = arbiter_ok(z, z32, z64)
    assert ok, (e_emu, e_ref)
    assert rel_l2(z, z32) <= 3 * e_ref + 1e-5


@pytest.mark.parametrize("name", ["cloth50_ragged", "rippled64_fragmented"])
def test_emulated_apply_with_the_top_walk_starting_at_level_1(name, emulator, synth, oracle_lib):
    """Small single-GPU meshes (at most 512 level-1 nodes): restrict_top takes over from level 1 instead of restrict_l1; same
    sums over the same groups, so z is bit-identical to the general launch sequence (rippled64 has 554 level-1 nodes: over the
    limit, nothing changes)."""
    mesh = CASES[name](synth)
    o32 = make_oracle(oracle_lib, mesh, "f")
    data = _input(o32, mesh, synth.residual(mesh.nv, 2))
    a = subprocess.run([emulator], input=data, capture_output=True, timeout=900, check=True).stdout
    b = subprocess.run([emulator], input=data, capture_output=True, timeout=900, check=True, env=dict(os.environ, MAS_EMU_TOP_FROM_L1="1")).stdout
    assert a == b and len(a) == 16 * mesh.nv


@pytest.mark.parametrize("name", ["cloth50_ragged", "cloud900_two_levels_multi_bank_top", "stacked2x24_ties"])
def test_emulated_apply_with_the_ancestor_walk(name, emulator, synth, oracle_lib):
    """Meshes whose whole level-0 solve runs beside the coarse chain: level-0 solve first, then the chain without prolong_sum,
    then add_coarse_walk (every vertex adds Z_1 + Z_2 + ... of its own ancestors, in prolong_sum's order): bit-identical z."""
    mesh = CASES[name](synth)
    o32 = make_oracle(oracle_lib, mesh, "f")
    data = _input(o32, mesh, synth.residual(mesh.nv, 2))
    a = subprocess.run([emulator], input=data, capture_output=True, timeout=900, check=True).stdout
    b = subprocess.run([emulator], input=data, capture_output=True, timeout=900, check=True, env=dict(os.environ, MAS_EMU_WALK="1")).stdout
    assert a == b and len(a) == 16 * mesh.nv


def test_emulated_apply_has_no_race_under_thread_sanitizer(tmp_path, synth, oracle_lib):
    exe, p = _build(tmp_path, extra=("-fsanitize=thread",))
    if p.returncode != 0:
        pytest.skip("ThreadSanitizer runtime not available: " + p.stderr[-200:])
    mesh = synth.cloth(64)                                     # three levels: restrict_fine / _l1, solve_coarse, prolong_sum, solve_fine
    o = make_oracle(oracle_lib, mesh, "f")
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=0")
    run = subprocess.run([exe], input=_input(o, mesh, synth.residual(mesh.nv)), capture_output=True, timeout=1800, env=env)
    assert run.returncode == 0, run.stderr[-500:]
    err = run.stderr.decode(errors="replace")
    assert err.count("WARNING: ThreadSanitizer: data race") == 0, err[:2000]

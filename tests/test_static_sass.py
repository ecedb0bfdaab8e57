"""Static checks of the machine code in libmas_b200.so (no GPU needed; cuobjdump ships with the toolkit): the library holds
sm_100a code only, the batched inversion (a20, LDLtInverse512 cpp:1347-1546) is on the tcgen05 tensor cores with its
accumulators in tensor memory, and the bandwidth-bound apply kernels (a21-a23, cpp:1548-1719) keep everything in registers.
Guards the properties profiles/r02_static_ptxas_sass.txt records against a later edit losing them silently."""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "preconditioner-for-cloth-and-deformable-body-simulation_b200", "libmas_b200.so")

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(LIB),
                                reason="needs cuobjdump and the built library")


@pytest.fixture(scope="module")
def sass():
    """kernel name (as in the source, template arguments dropped) -> Counter of SASS mnemonics (with their suffixes)"""
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    mangled = re.findall(r"Function : (\S+)", out)
    plain = subprocess.run(["c++filt"], input="\n".join(mangled), capture_output=True, text=True, check=True).stdout.split("\n")
    names = dict(zip(mangled, plain))
    hist, cur = collections.defaultdict(collections.Counter), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            k = re.search(r"mas::(?:\(anonymous namespace\)::)?(\w+_kernel)", names.get(m.group(1), ""))
            cur = k.group(1) if k else None
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            hist[cur][m.group(1)] += 1
    return hist


def count(h, prefix):
    return sum(n for op, n in h.items() if op.split(".")[0] == prefix)


def test_only_sm_100a_code_is_embedded():
    elf = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True, check=True).stdout
    ptx = subprocess.run(["cuobjdump", "-lptx", LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"\.(sm_\w+)\.cubin", elf))
    assert archs == {"sm_100a"}, archs
    assert "sm_" not in ptx.replace("sm_100a", ""), "PTX for another target would be a JIT fallback path"


def test_inversion_runs_on_tcgen05_with_tensor_memory(sass):
    for k in ("fine_assemble_invert_tc_kernel", "coarse_invert_tc_kernel"):
        h = sass[k]
        assert count(h, "UTCHMMA") >= 8, (k, "no tcgen05.mma in the SASS")
        assert count(h, "LDTM") >= 1 and count(h, "STTM") >= 1, (k, "accumulators do not live in tensor memory")
        assert count(h, "UTCBAR") >= 1, (k, "no tcgen05.commit")
        assert count(h, "HMMA") == 0 and count(h, "IMMA") == 0, (k, "legacy mma.sync found")


def test_streaming_apply_kernels_stay_in_registers(sass):
    for k in ("solve_fine_kernel", "restrict_fine_kernel", "add_coarse_kernel", "prolong_sum_kernel"):
        h = sass[k]
        assert h, k
        assert count(h, "STL") == 0 and count(h, "LDL") == 0, (k, "local-memory spills")
    h = sass["solve_fine_kernel"]
    wide = sum(n for op, n in h.items() if op.startswith("LDG") and ".128" in op)
    assert wide >= 37, "the packed inverses are read with 16-byte loads (37 per warp pass)"
    assert count(h, "FFMA") >= 200 and count(h, "SHFL") >= 90
    assert count(h, "LDS") == 0 and count(h, "STS") == 0, "level-0 solve uses no shared memory"


def test_every_kernel_of_the_launch_lists_is_in_the_library(sass):
    want = {"solve_fine_kernel", "solve_coarse_kernel", "restrict_fine_kernel", "restrict_l1_kernel", "restrict_top_kernel",
            "prolong_sum_kernel", "add_coarse_kernel", "gather_peers_kernel", "fine_assemble_invert_tc_kernel",
            "cross_bank_kernel", "collision_hessian_kernel", "connect_mask_l0_kernel", "close_components_kernel",
            "spmv_dot_kernel", "update_p_kernel", "pull_host_kernel", "copy_owned_kernel"}
    assert want <= set(sass), want - set(sass)

"""Static evidence that needs no GPU: ptxas resource usage of every kernel in libmas_b200.so and the memory / math
instruction mix of the hot kernels from the SASS (cuobjdump).  Writes profiles/<round>_static_ptxas_sass.txt.

    python tools/static_report.py [r01]
"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "preconditioner-for-cloth-and-deformable-body-simulation_b200")
HOT = ["solve_fine_kernel", "restrict_fine_kernel", "solve_coarse_kernel", "fine_assemble_invert_tc_kernel", "coarse_invert_tc_kernel", "fine_assemble_invert_kernel", "cross_bank_kernel",
       "spmv_dot_kernel", "pull_host_kernel"]


def demangle(name):
    out = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    return re.sub(r"\(anonymous namespace\)::", "", out).split("(")[0].replace("mas::", "")


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    lines = ["# ptxas -v (sm_100a), one row per kernel: registers / stack / spill stores+loads / static smem bytes", ""]
    pat = re.compile(r"Compiling entry function '(\S+)' for 'sm_100a'\nptxas info\s+: Function properties for \S+\n\s+(\d+) bytes stack "
                     r"frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers(?:, used \d+ barriers)?"
                     r"(?:, (\d+) bytes smem)?")
    for log in sorted(glob.glob(os.path.join(PKG, "csrc", "*.ptxas.log"))):
        unit = os.path.basename(log).replace(".ptxas.log", "")
        for m in pat.finditer(open(log).read()):
            name = demangle(m.group(1))
            if name.startswith("void cub::"):
                name = "cub::" + name.split("::")[2].split("<")[0]
            lines.append(f"{unit:13s} {name[:44]:44s} regs {int(m.group(5)):3d}  stack {int(m.group(2)):4d}  spill {int(m.group(3)):4d}+{int(m.group(4)):<4d} "
                         f"smem {m.group(6) or 0}")
    lines += ["", "# SASS instruction mix of the hot kernels (cuobjdump -sass libmas_b200.so): top mnemonics by count", ""]
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(PKG, "libmas_b200.so")], capture_output=True, text=True).stdout
    cur, hist = None, collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = demangle(m.group(1))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
        if m and cur:
            hist[cur][m.group(1)] += 1
    for k in HOT:
        if k not in hist:
            continue
        h = hist[k]
        total = sum(h.values())
        mem = {n: c for n, c in h.items() if re.match(r"(LDG|STG|LDS|STS|ATOM|RED|LDGSTS|SHFL|BAR|MATCH|REDUX)", n)}
        lines.append(f"{k}: {total} instructions")
        lines.append("  memory/sync: " + ", ".join(f"{n} x{c}" for n, c in sorted(mem.items(), key=lambda x: -x[1])))
        lines.append("  math       : " + ", ".join(f"{n} x{c}" for n, c in h.most_common(40)
                                                    if re.match(r"(FFMA|FMUL|FADD|DADD|DFMA|DMUL|MUFU|IMAD|HFMA2)", n)))
        # Blackwell tensor-core path: tcgen05.mma = UTC*MMA, tcgen05.ld / st = LDTM / STTM, tcgen05.alloc / commit = UTCATOMSWS / UTCBAR,
        # mbarrier = SYNCS, cvt.rna.tf32 = F2FP / I2FP-class conversions
        tens = {n: c for n, c in h.items() if re.match(r"(UTC|LDTM|STTM|SYNCS|F2FP|FENCE|UBLKCP|UTMA)", n)}
        if tens:
            lines.append("  tcgen05    : " + ", ".join(f"{n} x{c}" for n, c in sorted(tens.items(), key=lambda x: -x[1])))
        lines.append("")
    out = os.path.join(ROOT, "profiles", f"{tag}_static_ptxas_sass.txt")
    open(out, "w").write("\n".join(lines) + "\n")
    print(out)


if __name__ == "__main__":
    main()

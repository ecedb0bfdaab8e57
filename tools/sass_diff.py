"""Which kernels' machine code changed since a commit?  Compiles csrc/*.cu of the working tree and of <commit> to sm_100a
cubins (no GPU needed) and compares the SASS of every kernel by name.  Used to show that refactors / added experimental
variants leave the shipped kernels bit-identical.      python tools/sass_diff.py <commit>
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = "preconditioner-for-cloth-and-deformable-body-simulation_b200/csrc"
NVCC = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
        "--expt-relaxed-constexpr", "-cubin"]


def kernels(cubin):
    out = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            d = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout
            k = re.search(r"(\w+_kernel(?:<[^>]*>)?)", d)
            cur = k.group(1) if k and "cub::" not in d else None
            if cur:
                funcs[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and cur:
            # operand order of commutative integer adds and .reuse hints are scheduling noise
            funcs[cur].append(re.sub(r"\.reuse", "", m.group(1).strip()))
    return funcs


def main():
    commit = sys.argv[1]
    with tempfile.TemporaryDirectory() as tmp:
        old = os.path.join(tmp, "old", CSRC)
        os.makedirs(old)
        os.makedirs(os.path.join(tmp, "old", "include"))
        names = subprocess.run(["git", "-C", ROOT, "ls-tree", "--name-only", commit, CSRC + "/", "include/"], capture_output=True, text=True).stdout.split()
        for n in names:
            data = subprocess.run(["git", "-C", ROOT, "show", f"{commit}:{n}"], capture_output=True).stdout
            os.makedirs(os.path.dirname(os.path.join(tmp, "old", n)), exist_ok=True)
            open(os.path.join(tmp, "old", n), "wb").write(data)
        for unit in sorted(f for f in os.listdir(os.path.join(ROOT, CSRC)) if f.endswith(".cu")):
            res = {}
            for tag, base in (("old", os.path.join(tmp, "old", CSRC)), ("new", os.path.join(ROOT, CSRC))):
                src = os.path.join(base, unit)
                if not os.path.exists(src):
                    res[tag] = {}
                    continue
                cub = os.path.join(tmp, f"{tag}_{unit}.cubin")
                subprocess.run(NVCC + [src, "-o", cub], cwd=base, capture_output=True)
                res[tag] = kernels(cub) if os.path.exists(cub) else {}
            for k in sorted(res["new"]):
                ko = k if k in res["old"] else (k[:-3] if k.endswith("<0>") and k[:-3] in res["old"] else None)   # <0> = the shipped variant
                if ko is None:
                    state = "new"
                else:
                    a, b = res["old"][ko], res["new"][k]
                    norm = lambda x: re.sub(r"IADD3 (\S+), PT, PT, (\S+), (\S+), RZ", lambda m: "IADD3 " + m.group(1) + " " + " ".join(sorted([m.group(2), m.group(3)])), x)
                    state = "identical" if [norm(x) for x in a] == [norm(x) for x in b] else f"CHANGED ({len(a)} -> {len(b)} instructions)"
                print(f"{unit:18s} {k:44s} {len(res['new'][k]):6d}  {state}")


if __name__ == "__main__":
    main()

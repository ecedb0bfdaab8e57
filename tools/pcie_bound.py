"""Lower bound of the host-pointer apply (bench.py `e2e`): the two PCIe copies it cannot avoid.  Times a pinned-host
H2D and D2H copy of one xyzw vector field (16 B x nv) with CUDA events and prints one JSON line; the e2e time of
mas_apply(MAS_MEM_HOST) is  H2D + apply graph + D2H  because the coarse levels need all of r before any z is final.
The pinned buffer is allocated three times: with the process's default CPU affinity, from a CPU of the GPU's own NUMA
node, and from a CPU of another node (pinned pages live on the node of the allocating thread).

    python tools/pcie_bound.py [nv]
"""
import json
import os
import subprocess
import sys

import torch


def gpu_numa(index=0):
    """(numa node, local cpu list) of GPU `index` from sysfs; (None, None) when not visible."""
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(index)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        base = f"/sys/bus/pci/devices/{bus}"
        node = int(open(base + "/numa_node").read())
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        return node, cpus
    except Exception as e:                      # noqa: BLE001
        return None, str(e)


def measure(nv, dev):
    h = torch.zeros((nv, 4), dtype=torch.float32).pin_memory()
    d = torch.zeros((nv, 4), dtype=torch.float32, device=dev)
    out = {}
    for name, dst, src in (("h2d", d, h), ("d2h", h, d)):
        for _ in range(5):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        tot, n = 0.0, 20
        for _ in range(n):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dst.copy_(src, non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        out[name + "_ms"] = tot / n
        out[name + "_gbs"] = h.numel() * 4 / (tot / n * 1e-3) / 1e9
    return out


def mapped_view(h, dev):
    """CUDA-tensor view of a pinned host tensor (UVA: cudaHostAlloc memory has the same address on the device), so that an
    ordinary elementwise kernel moves the data with SM loads/stores over PCIe instead of the copy engine."""
    class _Wrap:
        pass
    w = _Wrap()
    w.__cuda_array_interface__ = {"shape": tuple(h.shape), "typestr": "<f4", "data": (h.data_ptr(), False), "version": 3,
                                  "strides": None}
    return torch.as_tensor(w, device=dev)


def measure_sm(nv, dev):
    h = torch.ones((nv, 4), dtype=torch.float32).pin_memory()
    d = torch.zeros((nv, 4), dtype=torch.float32, device=dev)
    m = mapped_view(h, dev)
    out = {}
    for name, fn in (("sm_pull", lambda: torch.mul(m, 1.0, out=d)), ("sm_push", lambda: torch.mul(d, 1.0, out=m))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        tot, n = 0.0, 10
        for _ in range(n):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        out[name + "_ms"] = tot / n
        out[name + "_gbs"] = nv * 16 / (tot / n * 1e-3) / 1e9
    out["pull_ok"] = bool((d == 1).all())
    return out


def main():
    nv = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
    dev = torch.device("cuda:0")
    torch.zeros(1, device=dev)
    allowed = os.sched_getaffinity(0)
    node, local = gpu_numa(0)
    out = {"nv": nv, "bytes_each_way": nv * 16, "gpu_numa_node": node, "allowed_cpus": len(allowed),
           "local_cpus": len(local) if isinstance(local, set) else local}
    out["default"] = measure(nv, dev)
    try:
        out["sm_zero_copy"] = measure_sm(nv, dev)
    except Exception as e:                      # noqa: BLE001
        out["sm_zero_copy"] = repr(e)
    if os.environ.get("PCIE_SKIP_AFFINITY"):
        local = None
    if isinstance(local, set):
        near, far = sorted(allowed & local), sorted(allowed - local)
        out["near_cpus_allowed"], out["far_cpus_allowed"] = len(near), len(far)
        for tag, cpus in (("gpu_local_node", near), ("other_node", far)):
            if not cpus:
                continue
            os.sched_setaffinity(0, set(cpus))
            out[tag] = measure(nv, dev)
            os.sched_setaffinity(0, allowed)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()

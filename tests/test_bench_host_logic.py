"""Host logic of bench.py that needs no GPU: the config dict both arms print (the driver compares them), the weak-scaling mesh
shapes, and the CPU reference leg run in a child process (with the fall-back to the plain-C port)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _args(**kw):
    base = dict(config=2, scaling="weak", gpus=1, proximity=False)
    base.update(kw)
    return argparse.Namespace(**base)


def test_config_dict_is_the_same_for_both_arms_and_names_the_workload():
    for world in (1, 2, 4, 8):
        cfg, units, nv = bench.workload_and_config(_args(gpus=world), world)
        assert units == float(world) and nv == 1048576 * world
        assert set(cfg) == {"workload", "nv", "units", "l2", "timing", "parallelism"}
        if world > 1:
            assert f"{world} x 1,048,576" in cfg["workload"] and "reference arm" in cfg["workload"]
    cfg, units, nv = bench.workload_and_config(_args(config=4, gpus=8), 8)
    assert units == 1.0 and nv == 4194304 and "strong scaling" in cfg["workload"]
    assert bench.weak_grid(1) == (1, 1) and bench.weak_grid(2) == (2, 1) and bench.weak_grid(4) == (2, 2) and bench.weak_grid(8) == (4, 2)


def test_cpu_leg_runs_in_a_child_process_and_returns_the_reference_z(pkg, synth, oracle_lib):
    mesh = synth.cloth(32)
    r = synth.residual(mesh.nv)
    rate, best, setup_ms, kind, z, note = bench.cpu_leg_isolated(mesh, r, 2, 1, 2)
    assert kind in ("reference", "port") and note is None and rate > 0 and best >= rate * 0.5 and setup_ms > 0
    o = oracle_lib.OraclePreconditioner("f")
    o.allocate(mesh)
    o.prepare()
    zo = o.apply(r)
    assert np.linalg.norm(z - zo) / np.linalg.norm(zo) < 1e-4
    # forced fall-back (what happens when the reference's own code crashes on an input)
    os.environ["MAS_BENCH_FORCE_PORT"] = "1"
    try:
        _, _, _, kind2, z2, _ = bench.cpu_leg_isolated(mesh, r, 1, 1, 1)
    finally:
        del os.environ["MAS_BENCH_FORCE_PORT"]
    assert kind2 == "port" and np.array_equal(z2, zo)

import importlib
import os
import sys

import pytest

# The in-process sharded tests run several contexts, each with its own apply-graph streams, that wait for each other ON THE
# DEVICE: every one of them needs a hardware work queue of its own (the default of 8 makes streams share queues, and a
# shard's publish can then sit behind another shard's spinning wait).  Must be set before CUDA initialises.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG_NAME = "preconditioner-for-cloth-and-deformable-body-simulation_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module(PKG_NAME)


@pytest.fixture(scope="session")
def synth(pkg):
    return pkg.synth


@pytest.fixture(scope="session")
def oracle_lib():
    """The plain-C restatement (oracle/mas_oracle.c); built on demand, test infrastructure only."""
    from oracle import oracle_binding as ob
    if not ob.available():
        import subprocess
        subprocess.run(["make"], cwd=os.path.join(ROOT, "oracle"), check=True, stdout=subprocess.DEVNULL)
    return ob


@pytest.fixture(scope="session")
def ref_lib():
    """The reference's own code compiled by oracle/build_ref.sh (prebuilt .so travels to the GPU box)."""
    from oracle import ref_binding as rb
    if not rb.available():
        pytest.skip("oracle/_ref/libmas_ref.so not built (needs /root/reference)")
    return rb

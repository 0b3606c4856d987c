import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure): built on demand, with the reference-compiled matrix_2d when present."""
    from oracle import pyoracle
    if not os.path.exists(pyoracle.LIB_PATH):
        pyoracle.build()
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def hostsim_path():
    """CPU-only build of the engine's host logic against plain-loop kernels (tests/hostsim)."""
    d = os.path.join(ROOT, "tests", "hostsim")
    subprocess.run(["make", "-s", "-C", d], check=True)
    return os.path.join(d, "_build", "libgadj_hostsim.so")


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library on a real device; never falls back to anything else."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from dynadjust_b200 import engine
    if not os.path.exists(engine.LIB_PATH):
        raise RuntimeError("dynadjust_b200/libgadj.so is missing on a GPU box: run __graft_entry__.build()")
    return engine.LIB_PATH

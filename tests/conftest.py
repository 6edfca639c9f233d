import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real CUDA device (run with -m gpu on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure): oracle/liboracle.so through ctypes."""
    import oracle_py
    oracle_py.build()
    return oracle_py


@pytest.fixture(scope="session")
def mirror():
    import np_mirror
    return np_mirror


@pytest.fixture(scope="session")
def emu_lib():
    """The CUDA sources compiled against the SIMT emulator (tests/emu) -- CPU-side check of the
    lane-parallel algorithms; never used by the product."""
    import idocp_b200
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "emu"), "-s"])
    return idocp_b200.Library(os.path.join(ROOT, "tests", "emu", "libidocp_b200_emu.so"))


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library (nvcc, sm_100a); fails loudly when it has not been built."""
    import idocp_b200
    return idocp_b200.default_library()


def make_states(batch, seed, scale_q=1.5, scale_v=0.5):
    rng = np.random.default_rng(seed)
    return rng.uniform(-scale_q, scale_q, (batch, 7)), rng.uniform(-scale_v, scale_v, (batch, 7))

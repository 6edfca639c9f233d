"""C++ host layer (include/idocp_b200/idocp_b200.hpp): the reference's class API over the C-ABI."""
import json
import os
import re
import subprocess

import pytest

from conftest import GOLDEN, ROOT

EXE = os.path.join(ROOT, "build", "unocp_benchmark")


def _build():
    import __graft_entry__ as g
    g.build_cuda()
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    lib = os.path.join(ROOT, "idocp_b200")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "unocp_benchmark.cpp"), "-L" + lib, "-lidocp_b200",
                           "-Wl,-rpath," + lib, "-o", EXE])


def test_example_compiles_and_fails_loudly_without_gpu():
    import torch
    _build()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    res = subprocess.run([EXE], capture_output=True, text=True)
    assert res.returncode != 0
    assert "no CUDA device" in res.stderr


@pytest.mark.gpu
def test_example_reproduces_golden_convergence():
    """examples/unocp_benchmark.cpp (twin of the reference example) prints the KKT history of the
    q = 2, v = 0 instance: must equal the committed golden vector digit for digit."""
    _build()
    out = subprocess.run([EXE, "3", "20"], capture_output=True, text=True, check=True).stdout
    kkt = [float(x) for x in re.findall(r"KKT error(?: after iteration \d+)? = (\S+)", out)]
    with open(os.path.join(GOLDEN, "unocp_golden.json")) as f:
        ref = json.load(f)["unocp_benchmark_reference_instance"]["kkt"]
    assert len(kkt) == 51
    assert kkt == ref
    assert "CPU time per update" in out

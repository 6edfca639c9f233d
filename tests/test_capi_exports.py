"""The nvcc-built product library loads and exports every symbol that include/idocp_b200.h
declares (no compute calls: there is no GPU in the build container)."""
import ctypes
import os
import re
import subprocess

import pytest

import idocp_b200 as I
from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "idocp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(idocp_b200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_bound_in_python():
    assert set(_declared()) == set(I.capi.EXPORTS)


def test_cuda_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    lib = g.build_cuda()
    L = ctypes.CDLL(lib)
    for name in _declared():
        assert hasattr(L, name), name
    L.idocp_b200_version.restype = ctypes.c_char_p
    assert b"sm_100a" in L.idocp_b200_version()


def test_cuda_library_contains_sm100a_code_only():
    import __graft_entry__ as g
    lib = g.build_cuda()
    out = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)


def test_no_gpu_no_fallback():
    """Without a CUDA device creation must fail loudly (never route to a CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = I.default_library()
    with pytest.raises(I.Idocp_b200Error, match="no CUDA device|CUDA"):
        I.UnOCPSolver(I.benchmark_problem(lib), 4)


def test_product_package_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "idocp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle_py" not in text and "liboracle" not in text and "np_mirror" not in text, f

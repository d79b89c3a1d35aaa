"""CPU tier: the C-ABI library builds, loads, exports every symbol include/bgpt_cuda.h
declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, have_gpu


def _declared():
    src = open(os.path.join(ROOT, "include", "bgpt_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bgpt_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(capi):
    L = capi.lib()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/bgpt_cuda.h but not exported"
    assert sorted(capi.SYMBOLS) == names


def test_no_torch_in_the_library(capi):
    """the shim is plain CUDA runtime: it must not pull torch / ATen into the process"""
    import subprocess
    out = subprocess.run(["ldd", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "c10" not in out


@pytest.mark.skipif(have_gpu(), reason="checks the behaviour WITHOUT a device")
def test_fails_loudly_without_gpu(capi):
    assert capi.device_count() <= 0
    hp = np.array([256, 2, 4, 64, 128, 64, 2], dtype=np.int32)
    assert not capi.lib().bgpt_cuda_model_create(hp, 0, 8)
    assert "no CPU path" in capi.last_error() or "CUDA" in capi.last_error()
    with pytest.raises(capi.BgptError):
        capi.op_norm(np.zeros((1, 64), np.float32))


def test_product_never_imports_oracle():
    """nothing under biogpt.cpp_b200/ may reference oracle/ (the oracle is a checker)"""
    pkg = os.path.join(ROOT, "biogpt.cpp_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp", ".hpp")) or f == "Makefile":
                s = open(os.path.join(dp, f), errors="ignore").read()
                assert "liboracle" not in s and "biogpt_oracle" not in s and "oracle/" not in s, os.path.join(dp, f)

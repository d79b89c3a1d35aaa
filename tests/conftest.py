"""Shared fixtures.  `-m "not gpu"` runs on the CPU-only builder box; `-m gpu` needs a B200."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

from _bootstrap import load_pkg  # noqa: E402

PKG = load_pkg()
gf = PKG.ggml_file

FTYPES = ["f32", "f16", "q4_0", "q4_1", "q5_0", "q5_1", "q8_0"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _build_checkers():
    """compile oracle/liboracle.so (and oracle/_ref when /root/reference exists) if missing"""
    import subprocess
    so = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference") and not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libbiogpt_ref.so")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True, stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def checkers():
    _build_checkers()
    import ref
    return ref


@pytest.fixture(scope="session")
def model_dir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("models"))


class ModelZoo:
    """synthetic `.bin` files, written on demand and cached for the session"""

    def __init__(self, root):
        self.root = root
        self._tensors = {}

    def tensors(self, size: str):
        if size not in self._tensors:
            hp = {"tiny": gf.TINY, "small": gf.SMALL, "base": gf.BASE, "narrow": gf.NARROW}[size]
            if size == "tiny":
                self._tensors[size] = golden_tiny_tensors()
            else:
                self._tensors[size] = gf.synth_tensors(hp, seed=1234)
        return self._tensors[size]

    def path(self, size: str, ftype: str) -> str:
        p = os.path.join(self.root, f"{size}-{ftype}.bin")
        if not os.path.exists(p):
            hp = {"tiny": gf.TINY, "small": gf.SMALL, "base": gf.BASE, "narrow": gf.NARROW}[size]
            gf.write_model(p, hp, self.tensors(size), gf.FTYPE_BY_NAME[ftype])
        return p


def golden_tiny_tensors():
    """the tiny model's f32 tensors as committed in tests/golden/tiny_model.npz"""
    z = np.load(os.path.join(ROOT, "tests", "golden", "tiny_model.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def zoo(model_dir):
    return ModelZoo(model_dir)


@pytest.fixture(scope="session")
def capi():
    import importlib
    return importlib.import_module("biogpt_cpp_b200.capi")


def have_gpu() -> bool:
    try:
        import importlib
        c = importlib.import_module("biogpt_cpp_b200.capi")
        return c.device_count() > 0
    except Exception:
        return False

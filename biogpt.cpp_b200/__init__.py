"""biogpt.cpp_b200 -- B200-native `biogpt_eval` hot path behind the reference's API.

Layout:
  csrc/      hand-written sm_100a CUDA kernels + the extern "C" shim (libbgpt_cuda.so)
  host/      C++ host side mirroring the reference's biogpt.h API (libbiogpt_b200.so)
  capi.py    ctypes view of include/bgpt_cuda.h (used by tests / bench.py, not by the product)
  ggml_file.py  `.bin` reader / writer / block codecs (fixture tooling)
"""
from . import ggml_file  # noqa: F401

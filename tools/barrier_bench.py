"""grid-barrier micro-benchmark: python tools/barrier_bench.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from _bootstrap import load_pkg
load_pkg()
import importlib
capi = importlib.import_module("biogpt_cpp_b200.capi")
names = ["counter+poll counter", "counter+flag(last arriver)", "per-CTA flags, 148 pollers/CTA", "red+poll counter(acquire)",
         "red+poll counter(relaxed)", "per-CTA flags, 1 warp polls", "two-level counters+flag"]
for v, n in enumerate(names):
    for wl in (0, 1):
        us = C.c_float(0)
        rc = capi.lib().bgpt_cuda_debug_barrier_bench(v, 2000, wl, C.byref(us))
        print(f"variant {v} ({n}){' + dependent L2 load' if wl else ''}: {us.value:.3f} us/barrier" if rc == 0 else capi.last_error())

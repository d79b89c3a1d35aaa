"""instruction-cache probe: cycles per instruction of a loop over KB kilobytes of straight-line code, all SMs at once"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from _bootstrap import load_pkg
load_pkg()
import importlib
capi = importlib.import_module("biogpt_cpp_b200.capi")
for nw in (1, 4, 16):
    for kb in (4, 8, 12, 16, 24, 32, 48, 64, 96, 128):
        cyc = C.c_float(0)
        rc = capi.tools_lib().bgpt_cuda_debug_icache_bench(kb, 50, nw, C.byref(cyc))
        n = kb * 64
        print(f"{nw:2d} warps/SM, loop body {kb:3d} KB ({n} instr): {cyc.value:9.0f} cycles/iteration = {cyc.value / n:.3f} cycles/instr/warp" if rc == 0 else capi.last_error())

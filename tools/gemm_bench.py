"""prompt-batch matmul benchmark: exact-order SIMT kernels vs the tcgen05 kernel.
   python tools/gemm_bench.py [--types q8_0,q4_0] [--n 8,64,128,512,1024]"""
import argparse, ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from _bootstrap import load_pkg
load_pkg()
import importlib
capi = importlib.import_module("biogpt_cpp_b200.capi")
ap = argparse.ArgumentParser()
ap.add_argument("--types", default="q8_0,q4_0")
ap.add_argument("--n", default="8,32,64,128,256,512,1024")
ap.add_argument("--shapes", default="1024x4096,4096x1024,1024x1024")   # KxROWS
a = ap.parse_args()
T = {"f32": 0, "f16": 1, "q4_0": 2, "q4_1": 3, "q5_0": 6, "q5_1": 7, "q8_0": 8}
for tn in a.types.split(","):
    for shp in a.shapes.split(","):
        k, rows = map(int, shp.split("x"))
        for n in map(int, a.n.split(",")):
            out = []
            for path in (0, 1):
                ms = C.c_float(0)
                iters = max(3, min(50, int(2e10 / (2.0 * k * rows * n + 1))))
                rc = capi.tools_lib().bgpt_cuda_debug_gemm_bench(T[tn], k, rows, n, iters, path, C.byref(ms))
                out.append((ms.value, 2.0 * k * rows * n / (ms.value * 1e-3) / 1e12) if rc == 0 else None)
            f = lambda o: f"{o[0] * 1e3:9.1f} us {o[1]:8.2f} TFLOP/s" if o else "      n/a"
            print(f"{tn:5s} K={k:5d} rows={rows:5d} n={n:5d}  simt-exact {f(out[0])}   tcgen05 {f(out[1])}")

"""the reference's front-end loop as written (examples/main/main.cpp:93-151: biogpt_eval with the whole logit row to the host, then
biogpt_sample_top_k_top_p), one token at a time through libbiogpt_b200.so.   python tools/main_loop_bench.py --ftype q4_0 --steps 256"""
import argparse, ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
ap = argparse.ArgumentParser()
ap.add_argument("--ftype", default="q4_0"); ap.add_argument("--steps", type=int, default=256); ap.add_argument("--n-past", type=int, default=0)
a = ap.parse_args()
H = C.CDLL(os.path.join(ROOT, "biogpt.cpp_b200", "host", "libbiogpt_b200.so"))
H.bgpt_host_open.restype = C.c_void_p
H.bgpt_host_open.argtypes = [C.c_char_p, C.c_int]
H.bgpt_host_close.argtypes = [C.c_void_p]
H.bgpt_host_main_loop.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_uint32, C.c_void_p, C.POINTER(C.c_double)]
hs = H.bgpt_host_open(bench.model_path(a.ftype).encode(), 8)
ids = np.zeros(a.n_past + a.steps, np.int32); ws = C.c_double(0)
if a.n_past:
    H.bgpt_host_main_loop(hs, 2, 0, a.n_past, 40, 0.9, 0.8, 1, ids.ctypes.data, C.byref(ws))
for rep in range(3):
    rc = H.bgpt_host_main_loop(hs, 2, a.n_past, a.steps, 40, 0.9, 0.8, 1, ids.ctypes.data, C.byref(ws))
    assert rc == 0, rc
    print(f"{a.ftype} main.cpp loop (biogpt_eval + biogpt_sample_top_k_top_p), n_past {a.n_past}..{a.n_past + a.steps - 1}: {ws.value / a.steps * 1e6:.1f} us per token")
H.bgpt_host_close(hs)

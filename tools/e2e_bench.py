"""per-token cost of the reference-facing call: bgpt_cuda_eval_topk with host buffers (what biogpt_eval_sample calls), wall clock
against the device time of the same steps.   python tools/e2e_bench.py --ftype q4_0 --steps 256 --n-past 0"""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, importlib
ap = argparse.ArgumentParser()
ap.add_argument("--ftype", default="q4_0"); ap.add_argument("--steps", type=int, default=256); ap.add_argument("--n-past", type=int, default=0)
ap.add_argument("--k", type=int, default=40)
a = ap.parse_args()
capi = importlib.import_module("biogpt_cpp_b200.capi")
M = capi.Model.load(bench.model_path(a.ftype))
tok = np.array([2], dtype=np.int32)
if a.n_past:
    M.decode_greedy(2, 0, a.n_past)
for rep in range(3):
    t0 = time.perf_counter()
    for p in range(a.n_past, a.n_past + a.steps):
        vals, ids, exact, full = M.eval_topk(tok, p, a.k)
        tok[0] = int(ids[0]) if exact else int(np.argmax(full))
    wall = (time.perf_counter() - t0) / a.steps * 1e6
# a separate pass: CUDA-event time of each call (first launch .. result packet written), i.e. what the GPU spends per call
ev = 0.0
t0 = time.perf_counter()
for p in range(a.n_past, a.n_past + a.steps):
    vals, ids, exact, full = M.eval_topk(tok, p, a.k)
    ev += M.last_eval_ms
    tok[0] = int(ids[0]) if exact else int(np.argmax(full))
wall2 = (time.perf_counter() - t0) / a.steps * 1e6
print(f"   CUDA-event time per call {ev * 1e3 / a.steps:.1f} us (pass wall {wall2:.1f} us)")
# pure call cost without numpy post-processing
L = capi.lib(); import ctypes as C
vals = np.zeros(a.k, np.float32); idsb = np.zeros(a.k, np.int32); n_out = C.c_int(0); ex = C.c_int(0)
t0 = time.perf_counter()
for p in range(a.n_past, a.n_past + a.steps):
    L.bgpt_cuda_eval_topk(M.h, tok, 1, p, a.k, vals, idsb, C.byref(n_out), C.byref(ex), None)
    tok[0] = idsb[0]
wall3 = (time.perf_counter() - t0) / a.steps * 1e6
print(f"   raw ctypes call loop {wall3:.1f} us per token")
# the same loop in C++ (host/host_capi.cpp: bgpt_host_sampling_loop) -- what a C++ caller such as examples/main pays
H = C.CDLL(os.path.join(ROOT, "biogpt.cpp_b200", "host", "libbiogpt_b200.so"))
H.bgpt_host_sampling_loop.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_double)]
out_ids = np.zeros(a.steps, np.int32); ws = C.c_double(0)
for rep in range(2):
    rc = H.bgpt_host_sampling_loop(M.h, 2, a.n_past, a.steps, a.k, out_ids.ctypes.data, C.byref(ws))
    assert rc == 0, rc
print(f"   C++ loop {ws.value / a.steps * 1e6:.1f} us per token")
ids_dev, ms = M.decode_greedy(2, a.n_past, a.steps) if a.n_past == 0 else (None, 0.0)
dev = ms * 1e3 / a.steps if ms else float("nan")
print(f"{a.ftype} eval_topk k={a.k}: {wall:.1f} us per token wall (n_past {a.n_past}..{a.n_past + a.steps - 1}); device-resident greedy loop {dev:.1f} us per token; "
      f"overhead {wall - dev:.1f} us = {100 * (wall - dev) / wall:.1f} %")
M.close()

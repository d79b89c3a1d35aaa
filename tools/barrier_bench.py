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
        rc = capi.tools_lib().bgpt_cuda_debug_barrier_bench(v, 2000, wl, C.byref(us))
        print(f"variant {v} ({n}){' + dependent L2 load' if wl else ''}: {us.value:.3f} us/barrier" if rc == 0 else capi.last_error())

xn = ["exchange: ld.volatile, 8 replicas, 512 pollers (generation-4 kernel)", "exchange: ld.relaxed.gpu, 8 replicas, 512 pollers", "exchange: ld.acquire.gpu, 8 replicas, 512 pollers",
      "exchange: ld.volatile, 1 replica", "exchange: ld.volatile, 2 replicas", "exchange: ld.volatile, 8 replicas, 128 pollers x 4", "exchange: ld.relaxed.gpu, 1 replica",
      "exchange: ld.relaxed.gpu, 8 replicas, 128 pollers x 4"]
for i, n in enumerate(xn):
    for sl in (0, 100):
        us = C.c_float(0)
        rc = capi.tools_lib().bgpt_cuda_debug_barrier_bench(7 + i, 2000, sl, C.byref(us))
        print(f"variant {7 + i} ({n}){', nanosleep(100) back-off' if sl else ''}: {us.value:.3f} us/exchange" if rc == 0 else capi.last_error())
for i, n in enumerate(["ld.volatile", "ld.relaxed.gpu", "ld.acquire.gpu"]):
    for peer in (1, 2, 75, 147):
        us = C.c_float(0)
        rc = capi.tools_lib().bgpt_cuda_debug_barrier_bench(15 + i, 2000, peer, C.byref(us))
        print(f"ping-pong CTA 0 <-> CTA {peer} ({n}): {us.value * 500:.0f} ns one way" if rc == 0 else capi.last_error())

for v, n in [(18, "exchange: ld.volatile, 16 replicas"), (19, "exchange: ld.volatile, 32 replicas"), (20, "exchange: ld.global.cg, 8 replicas"),
             (21, "packed exchange {3 x f32, tag} 16-byte units, 8 replicas, 384 pollers"), (22, "packed exchange, 16 replicas"), (23, "packed exchange, 1 replica")]:
    us = C.c_float(0)
    rc = capi.tools_lib().bgpt_cuda_debug_barrier_bench(v, 2000, 0, C.byref(us))
    print(f"variant {v} ({n}): {us.value:.3f} us/exchange" if rc == 0 else capi.last_error())

for v, n in [(24, "producer: lane 28 of warps < rows, 8 sequential stores"), (25, "producer: threads 8*row, 8 sequential stores"),
             (26, "producer: gather + sync, warp 0, 4 replicas x rows per store, 2 stores"), (27, "producer: gather + sync, warp 0 lanes < rows, 8 sequential stores"),
             (28, "producer: gather + sync, warp r writes replica r")]:
    for rep in range(2):
        us = C.c_float(0)
        rc = capi.tools_lib().bgpt_cuda_debug_barrier_bench(v, 2000, 0, C.byref(us))
        print(f"variant {v} ({n}): {us.value:.3f} us/exchange" if rc == 0 else capi.last_error())

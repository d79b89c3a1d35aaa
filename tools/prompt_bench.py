"""prompt processing: 1024 tokens evaluated in chunks of N (the reference's n_batch), BioGPT-base.
   python tools/prompt_bench.py --ftype q8_0 --n 8,64,256,1024
FLOPs per eval: 2*N*301,989,888 + 2*43,401,216 + 98,304*N*(n_past+N)."""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import importlib
ap = argparse.ArgumentParser()
ap.add_argument("--ftype", default="q8_0")
ap.add_argument("--n", default="8,32,64,128,256,512,1024")
ap.add_argument("--total", type=int, default=1024)
a = ap.parse_args()
capi = importlib.import_module("biogpt_cpp_b200.capi")
gf = bench.gf
M = capi.Model.load(bench.model_path(a.ftype), max_batch=1024)
toks = gf.synth_tokens(a.total, gf.BASE.n_vocab, seed=5)
for N in map(int, a.n.split(",")):
    for rep in range(2):
        t_dev = 0.0; flops = 0.0
        t0 = time.perf_counter()
        for p in range(0, a.total, N):
            M.eval(toks[p:p + N], p)
            t_dev += M.last_eval_ms
            flops += 2.0 * N * 301989888 + 2 * 43401216 + 98304.0 * N * (p + N)
        wall = time.perf_counter() - t0
    print(f"{a.ftype} n_batch={N:5d}: {a.total} prompt tokens in {t_dev:9.2f} ms device ({wall * 1e3:9.2f} ms wall) -> "
          f"{a.total / (t_dev * 1e-3):9.0f} tok/s, {flops / (t_dev * 1e-3) / 1e12:7.2f} TFLOP/s")
M.close()

"""one prompt eval of N tokens for ncu: python tools/profile_prompt.py --ftype q8_0 --n 1024"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, importlib
ap = argparse.ArgumentParser()
ap.add_argument("--ftype", default="q8_0"); ap.add_argument("--n", type=int, default=1024); ap.add_argument("--reps", type=int, default=1)
a = ap.parse_args()
capi = importlib.import_module("biogpt_cpp_b200.capi")
M = capi.Model.load(bench.model_path(a.ftype), max_batch=a.n)
toks = bench.gf.synth_tokens(a.n, bench.gf.BASE.n_vocab, seed=5)
for _ in range(a.reps):
    M.eval(toks, 0)
print(f"{a.ftype} n={a.n}: {M.last_eval_ms:.3f} ms")
M.close()

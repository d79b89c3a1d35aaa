"""prompt evals for ncu.
   python tools/profile_prompt.py --ftype q8_0 --n 1024                 one eval of N tokens at n_past 0 (tcgen05 path for N >= 32)
   python tools/profile_prompt.py --ftype q8_0 --n 8 --n-past 512       fill 512 positions with one big eval, then `--reps` evals of 8
                                                                        rows (the reference's n_batch; fused skinny-batch schedule)"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, importlib
ap = argparse.ArgumentParser()
ap.add_argument("--ftype", default="q8_0"); ap.add_argument("--n", type=int, default=1024); ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--n-past", type=int, default=0)
a = ap.parse_args()
capi = importlib.import_module("biogpt_cpp_b200.capi")
M = capi.Model.load(bench.model_path(a.ftype), max_batch=max(a.n, a.n_past, 8))
toks = bench.gf.synth_tokens(1024, bench.gf.BASE.n_vocab, seed=5)
pos = 0
if a.n_past:
    M.eval(toks[:a.n_past], 0); pos = a.n_past
for _ in range(a.reps):
    M.eval(toks[pos:pos + a.n], pos); pos += a.n
    print(f"{a.ftype} n={a.n} n_past={pos - a.n}: {M.last_eval_ms:.3f} ms, {M.launch_count} launches so far")
M.close()

"""Short decode run for ncu: loads the synthetic BioGPT-base model and decodes a few tokens at a
chosen position (the KV cache before it is whatever is resident -- timing only).
  python tools/profile_decode.py --ftype q4_0 --n-past 511 --steps 4 --warm 2
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import importlib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ftype", default="q4_0")
ap.add_argument("--n-past", type=int, default=511)
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--warm", type=int, default=2)
a = ap.parse_args()
capi = importlib.import_module("biogpt_cpp_b200.capi")
M = capi.Model.load(bench.model_path(a.ftype))
if a.warm:
    M.decode_greedy(2, a.n_past, a.warm)
ids, ms = M.decode_greedy(2, a.n_past, a.steps)
print(f"{a.ftype} n_past={a.n_past}: {a.steps} tokens in {ms:.3f} ms -> {ms / a.steps * 1e3:.1f} us/token, launches={M.launch_count}")
M.close()

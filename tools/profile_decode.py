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
prof = M.read_prof()
if prof is not None:
    import numpy as np
    names = ["LN0+qkv", "attention", "out_proj", "LN1+fc1", "fc2"]
    L = prof.shape[0] - 1
    P = prof[:L].astype(np.float64)
    nxt = np.empty((L, 5)); nxt[:, :4] = P[:, 1:, 0]; nxt[:-1, 4] = P[1:, 0, 0]; nxt[-1, 4] = prof[L, 0, 0]
    print("per-phase cycles of CTA 0 (mean over layers; ~1.9 cycles/ns): prologue | matmul/attn | barrier wait | total")
    for i, nme in enumerate(names):
        pro = (P[:, i, 1] - P[:, i, 0]).mean() if i != 1 else 0.0
        body = (P[:, i, 2] - P[:, i, 1]).mean() if i != 1 else (P[:, i, 2] - P[:, i, 0]).mean()
        bar = (nxt[:, i] - P[:, i, 2]).mean()
        print(f"  {nme:10s} {pro:9.0f} {body:9.0f} {bar:9.0f} {pro + body + bar:9.0f}")
    print("  LN0 prologue split (cycles): prefetch-issue %.0f | x load %.0f | LayerNorm %.0f | quantise %.0f" % (
        (P[1:, 0, 5] - P[1:, 0, 0]).mean(), (P[1:, 0, 3] - P[1:, 0, 5]).mean(), (P[:, 0, 4] - P[:, 0, 3]).mean(), (P[:, 0, 1] - P[:, 0, 4]).mean()))
    print("  LN1 prologue split (cycles): x1 load %.0f | LayerNorm %.0f | quantise %.0f" % (
        (P[:, 3, 3] - P[:, 3, 0]).mean(), (P[:, 3, 4] - P[:, 3, 3]).mean(), (P[:, 3, 1] - P[:, 3, 4]).mean()))
    print(f"  layer total {(nxt[:, 4] - P[:, 0, 0]).mean():9.0f} cycles;  lm_head: prologue {prof[L,0,1]-prof[L,0,0]} matmul {prof[L,0,2]-prof[L,0,1]}; whole kernel {prof[L,0,2]-prof[0,0,0]}")
    if os.environ.get("M4_EXP"):
        for i, nme in enumerate(names):
            if i == 1: continue
            print(f"  {nme:10s} phase A first run {(P[:, i, 4] - P[:, i, 1]).mean():8.0f} cycles, second run {(P[:, i, 5] - P[:, i, 4]).mean():8.0f}")
M.close()

"""Timeline of one launch of the persistent multi-row kernel (csrc/bgpt_rows.cuh): clock64 stamps of CTA 0 at the end of
every stage (BGPT_MEGA_PROF=1).  Prints the mean duration of each stage over the layers and the per-layer total.
   BGPT_MEGA_PROF=1 python tools/trace_rows.py --ftype q5_1 --rows 8 --mode streams --n-past 511"""
import argparse, os, sys
import numpy as np
os.environ.setdefault("BGPT_MEGA_PROF", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import importlib
ap = argparse.ArgumentParser()
ap.add_argument("--ftype", default="q5_1")
ap.add_argument("--rows", type=int, default=8)
ap.add_argument("--mode", default="streams", choices=["streams", "prompt"])
ap.add_argument("--n-past", type=int, default=511)
a = ap.parse_args()
capi = importlib.import_module("biogpt_cpp_b200.capi")
gf = bench.gf
M = capi.Model.load(bench.model_path(a.ftype), max_batch=8)
M.set_batch_path(2)                                  # the multi-row kernel is opt-in
tok = gf.synth_tokens(a.rows, gf.BASE.n_vocab, seed=9).astype(np.int32)
if a.mode == "streams":
    M.set_streams(a.rows)
    for i in range(4):
        M.eval_streams(tok, a.n_past - 3 + i)
else:
    for i in range(4):
        M.eval(tok, a.n_past)
ms = M.last_eval_ms
t = M.read_rows_trace()
assert t is not None, "no trace: BGPT_MEGA_PROF=1 must be set before the model is loaded"
L = t.shape[0] - 1
ghz = 1.965
start = t[L][7]
names = ["A ln0", "B qkv", "C attn", "D out_proj", "E ln1", "F fc1", "G fc2"]
print(f"{a.ftype} {a.mode} rows={a.rows} n_past={a.n_past}: eval {ms * 1e3:.1f} us (event), kernel {(t[L][1] - start) / ghz / 1e3:.1f} us (CTA 0 clock, {ghz} GHz assumed)")
prev = np.concatenate([[start], t[:L - 1, 6]])
dur = np.zeros((L, 7))
for l in range(L):
    p = prev[l]
    for s in range(7):
        dur[l, s] = (t[l][s] - p) / ghz
        p = t[l][s]
for s in range(7):
    print(f"  {names[s]:<11} mean {dur[1:, s].mean():8.0f} ns   (layer 0: {dur[0, s]:8.0f}, min {dur[1:, s].min():8.0f}, max {dur[1:, s].max():8.0f})")
print(f"  layer       mean {dur[1:].sum(axis=1).mean():8.0f} ns;  final LN + lm_head {(t[L][1] - t[L - 1][6]) / ghz:8.0f} ns")
M.close()

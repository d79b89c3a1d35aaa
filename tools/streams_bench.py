"""BASELINE.json configs[3]: S independent streams decoded in lock-step on one GPU (each weight read is shared by the S
streams; every stream has its own F32 KV cache), fused skinny-batch schedule (BGPT_BATCH_PATH=0: per-operator schedule).  Host argmax per stream (greedy).
   python tools/streams_bench.py --ftype q5_1 --streams 8 --steps 64 [--n-past 0]"""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import importlib
ap = argparse.ArgumentParser()
ap.add_argument("--ftype", default="q5_1")
ap.add_argument("--streams", type=int, default=8)
ap.add_argument("--steps", type=int, default=64)
ap.add_argument("--n-past", type=int, default=0)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
capi = importlib.import_module("biogpt_cpp_b200.capi")
gf = bench.gf
M = capi.Model.load(bench.model_path(a.ftype), max_batch=max(8, a.streams))
M.set_streams(a.streams)
tok = gf.synth_tokens(a.streams, gf.BASE.n_vocab, seed=9).astype(np.int32)
for warm in range(a.reps):
    t_dev = 0.0
    cur = tok.copy()
    t0 = time.perf_counter()
    for i in range(a.steps):
        logits = M.eval_streams(cur, a.n_past + i)
        t_dev += M.last_eval_ms
        cur = np.argmax(logits, axis=1).astype(np.int32)
    wall = time.perf_counter() - t0
W = 345_391_104 * {"f16": 64, "q4_0": 18, "q4_1": 20, "q5_0": 22, "q5_1": 24, "q8_0": 34}[a.ftype] // 32
kv = sum(196_608 * (a.n_past + i + 1) for i in range(a.steps)) / a.steps
B = W + a.streams * (kv + 1_286_144 // a.streams + 196_608 + 169_536)
print(f"{a.ftype} {a.streams} lock-step streams, n_past {a.n_past}..{a.n_past + a.steps - 1}: {t_dev / a.steps * 1e3:.1f} us/step device "
      f"({wall / a.steps * 1e6:.1f} us wall) -> {a.streams * a.steps / (t_dev * 1e-3):.0f} tok/s device, {a.streams * a.steps / wall:.0f} tok/s wall; "
      f"algorithmic {B / 1e6:.1f} MB/step -> {B / (t_dev / a.steps * 1e-3) / 1e9:.1f} GB/s")
M.close()

"""chained launches of bgpt_cuda_eval_topk under a short patience: BGPT_CHAIN_WAIT_US=5 makes nearly every queued kernel give up before its
token comes (the call then serves the position with a fresh launch); the sampled ids must equal the device-resident greedy loop's.
   for w in 5 50 400 2000; do BGPT_CHAIN_WAIT_US=$w python tools/chain_stress.py; done"""
import os, sys, time, importlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
capi = importlib.import_module("biogpt_cpp_b200.capi")
M = capi.Model.load(bench.model_path("q4_0"))
ids_dev, _ = M.decode_greedy(2, 0, 300)
tok = np.array([2], np.int32); out = []
t0 = time.perf_counter()
for p in range(300):
    vals, ids, exact, full = M.eval_topk(tok, p, 40)
    nxt = int(ids[0]) if exact else int(np.argmax(full))
    out.append(nxt); tok[0] = nxt
    if p % 7 == 3: time.sleep(0.0003)
print("wait_us", os.environ.get("BGPT_CHAIN_WAIT_US"), "ids equal device loop:", out == ids_dev.tolist(), f"{(time.perf_counter()-t0)/300*1e6:.1f} us/token")
M.close()

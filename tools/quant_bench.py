"""device weight quantiser (f32 -> Qx): microseconds and GB/s for the BioGPT-base weight set
   python tools/quant_bench.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from _bootstrap import load_pkg
gf = load_pkg().ggml_file
import importlib
capi = importlib.import_module("biogpt_cpp_b200.capi")
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6550.0)
n = 345_391_104                      # all matmul weights of BioGPT-base
for name, t in (("q4_0", 2), ("q4_1", 3), ("q5_0", 6), ("q5_1", 7), ("q8_0", 8)):
    us = capi.quantize_bench(t, n, 5)
    b = n * 4 + gf.row_bytes(t, n)
    print(f"{name}: {n} weights in {us:9.1f} us -> {b / us / 1e3:7.1f} GB/s algorithmic (read 4 B + write {gf.TYPE_SIZE[t] / 32:.4f} B per weight) = {b / us / 1e3 / peak * 100:5.1f} % of {peak:.0f} GB/s")

"""Workload for compute-sanitizer (tools/gpu_r2.sh sanitize): every schedule once on small synthetic models through the C ABI --
prompt batches (per-operator, skinny-batch, bit-exact tcgen05), single-token steps on the three persistent-kernel generations,
lock-step streams, the device top-k.  --quick: the decode kernels only (racecheck is slow)."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from _bootstrap import load_pkg  # noqa: E402

gf = load_pkg().ggml_file
import importlib  # noqa: E402

capi = importlib.import_module("biogpt_cpp_b200.capi")
quick = "--quick" in sys.argv
tmp = tempfile.mkdtemp()
WIDE = gf.HParams(**{**gf.NARROW.__dict__, "n_vocab": 4500})        # vocabulary >= 4096: the two- / three-launch device top-k
jobs = [("narrow", gf.NARROW, "q4_0"), ("narrow", gf.NARROW, "f16")] if quick else [
    ("narrow", gf.NARROW, "q4_0"), ("narrow", gf.NARROW, "q5_1"), ("narrow", gf.NARROW, "f16"), ("wide", WIDE, "q8_0"), ("wide", WIDE, "q5_0"), ("small", gf.SMALL, "f16"), ("tiny", gf.TINY, "q8_0")]
for name, hp, ft in jobs:
    path = os.path.join(tmp, f"{name}-{ft}.bin")
    gf.write_model(path, hp, gf.synth_tensors(hp, seed=1234), gf.FTYPE_BY_NAME[ft])
    M = capi.Model.load(path, max_batch=128)
    toks = gf.synth_tokens(hp.n_positions, hp.n_vocab, seed=3)
    pos = 0
    if not quick:
        for n in (8, 5, 16):
            M.eval(toks[pos:pos + n], pos); pos += n
        if hp.n_positions >= 256:
            M.eval(toks[pos:pos + 128], pos); pos += 128       # warp-specialised bit-exact tcgen05 matmul (quantised) / per-operator (f16)
            if ft == "f16" and name == "narrow":
                M.set_f16_tc_min_rows(32)
                M.eval(toks[pos:pos + 130], pos); pos += 130   # opt-in K-accumulating tcgen05 matmul, ragged token tile
                M.set_f16_tc_min_rows(0)
            elif name == "narrow":
                M.set_tcw(0); M.eval(toks[pos:pos + 128], pos); pos += 128; M.set_tcw(1)     # the single-stage form
    for path_id in ((1, 3, 2) if name in ("narrow", "wide") else (1,)):
        M.set_decode_path(path_id)
        for i in range(3):
            M.eval(toks[pos + i:pos + i + 1], pos + i)
    M.set_decode_path(1)
    M.decode_greedy(2, pos, 4)
    # the sampler's form of the generation-5 kernel: selection as the kernel's tail, the next position's launch queued and fed its token
    # through mapped host memory (chained), a withdrawal (the jump back), and the unchained form
    for p2 in (pos, pos + 1, pos + 2, pos):
        M.eval_topk(toks[p2:p2 + 1], p2, 40)
    M.set_chain(0); M.eval_topk(toks[pos:pos + 1], pos, 5); M.set_chain(-1)
    if name in ("narrow", "wide"):
        M.set_decode_path(3); M.eval_topk(toks[pos:pos + 1], pos, 40); M.set_decode_path(1)     # generation 4: selection in launches behind the kernel
    if not quick:
        M.eval_topk(toks[pos:pos + 4], pos, 40)                # after a prompt batch: slices -> groups -> one CTA
        M.set_streams(3)
        M.eval_streams(toks[:3], 0)
        M.decode_greedy_streams(toks[:3], 0, 3)
    M.close()
    print(f"sanitize_run: {name}/{ft} done")

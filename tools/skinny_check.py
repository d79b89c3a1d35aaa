"""Localise a mismatch between the fused skinny-batch schedule (csrc/bgpt_skinny.cuh) and the per-operator schedule:
run the same eval on both and compare the arena buffers the LAST layer left behind.
   python tools/skinny_check.py [--ftypes q4_0,q5_1] [--n 8] [--pos 0,40]
which buffer differs first names the kernel: q -> LayerNorm0 + q,k,v; x1 -> attention or out_proj; act_ff -> LayerNorm1 + fc1 +
GELU + quantise; x -> fc2; logits -> final LayerNorm + lm_head."""
import argparse, os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from _bootstrap import load_pkg
PKG = load_pkg()
gf = PKG.ggml_file
import importlib
capi = importlib.import_module("biogpt_cpp_b200.capi")
ap = argparse.ArgumentParser()
ap.add_argument("--ftypes", default="q4_0,q4_1,q5_0,q5_1,q8_0")
ap.add_argument("--n", type=int, default=8)
ap.add_argument("--pos", default="0,8,40")
a = ap.parse_args()
hp = gf.NARROW
tens = gf.synth_tensors(hp, seed=1234)
tmp = tempfile.mkdtemp()
toks = gf.synth_tokens(hp.n_positions, hp.n_vocab, seed=55)
names = {0: "x (fc2 out)", 1: "x1 (out_proj out)", 2: "q", 4: "act_ff (fc1 out)"}
bad_total = 0
for ft in a.ftypes.split(","):
    p = os.path.join(tmp, f"narrow-{ft}.bin")
    gf.write_model(p, hp, tens, gf.FTYPE_BY_NAME[ft])
    M = capi.Model.load(p, max_batch=32)
    for mode in ("prompt", "streams"):
        if mode == "streams":
            M.set_streams(a.n)
        for pos in map(int, a.pos.split(",")):
            res = {}
            for path in (1, 0):
                M.set_batch_path(path)
                if mode == "prompt":
                    lg = M.eval(toks[pos:pos + a.n], pos)
                else:
                    lg = M.eval_streams(toks[pos:pos + a.n], pos)
                res[path] = {"logits": np.ascontiguousarray(lg).view(np.uint8).ravel()}
                for w in names:
                    res[path][names[w]] = M.read_buffer(w, a.n)
            line = []
            for k in ["q", "x1 (out_proj out)", "act_ff (fc1 out)", "x (fc2 out)", "logits"]:
                f, o = res[1][k], res[0][k]
                if k.startswith("act_ff"):                      # the fused schedule does not fill the offsets of formats that have none
                    pass
                nb = int(np.count_nonzero(f != o))
                bad_total += nb
                line.append(f"{k}: {nb}/{f.size} bytes differ")
            print(f"{ft} {mode} n={a.n} n_past={pos} fused={M.batch_path(a.n) == 0 and 'see path flag' or 'on'}: " + "; ".join(line), flush=True)
    M.close()
print("TOTAL differing bytes:", bad_total)

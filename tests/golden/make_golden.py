"""Generates the committed fixtures in this directory FROM THE UNMODIFIED REFERENCE
(oracle/_ref/libbiogpt_ref.so, built from /root/reference by oracle/Makefile).  Run in the
builder container only:   python tests/golden/make_golden.py

  tiny_model.npz       the f32 tensors of the tiny synthetic model (seed 1234); every test
                       rebuilds the seven `.bin` files from these, so fixtures do not depend on
                       numpy's random stream staying stable
  tiny_logits.npz      per ftype: sha256 of the `.bin`, and the reference's last-row logits for
                       a fixed schedule of (tokens, n_past) evals (prompt batches of 5 and 3
                       tokens, then single-token decode steps -- un-masked attention makes the
                       chunking part of the contract)
  quantize_ref.npz     reference quantize-tool output hashes for the tiny model (pins the numpy
                       block quantisers in ggml_file.py) and block-codec known answers on the
                       ggml test signal 0.1 + 2 cos(i) (ggml/tests/test-quantize-fns.cpp:26-30)
  tables.npz           the reference's fp16 GELU / exp tables probed through ggml_gelu /
                       ggml_soft_max on all 65536 fp16 inputs
"""
import hashlib
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from _bootstrap import load_pkg  # noqa: E402

gf = load_pkg().ggml_file
import ref  # noqa: E402

SCHEDULE = [(5, 0), (3, 5), (1, 8), (1, 9), (1, 10), (1, 11), (2, 12), (1, 14)]  # (n tokens, n_past)


def schedule_tokens(n_vocab):
    return gf.synth_tokens(sum(n for n, _ in SCHEDULE), n_vocab, seed=7)


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def main():
    assert ref.have_ref(), "build oracle/_ref first (make -C oracle)"
    hp = gf.TINY
    tensors = gf.synth_tensors(hp, seed=1234)
    np.savez_compressed(os.path.join(HERE, "tiny_model.npz"), **tensors)
    tmp = tempfile.mkdtemp()
    toks = schedule_tokens(hp.n_vocab)
    out = {"tokens": toks, "schedule": np.array(SCHEDULE, dtype=np.int32)}
    qout = {}
    f32_path = os.path.join(tmp, "tiny-f32.bin")
    gf.write_model(f32_path, hp, tensors, 0)
    for name, ft in gf.FTYPE_BY_NAME.items():
        p = os.path.join(tmp, f"tiny-{name}.bin")
        gf.write_model(p, hp, tensors, ft)
        out[f"sha_{name}"] = np.frombuffer(bytes.fromhex(sha(p)), dtype=np.uint8)
        if ft >= 2:  # reference quantize tool on the f32 file must give the same bytes
            q = os.path.join(tmp, f"tiny-{name}-refq.bin")
            ref.ref_quantize_file(f32_path, q, ft)
            qout[f"sha_{name}"] = np.frombuffer(bytes.fromhex(sha(q)), dtype=np.uint8)
        R = ref.Ref(p)
        logits = []
        pos = 0
        for n, n_past in SCHEDULE:
            logits.append(R.eval(toks[pos:pos + n], n_past))
            pos += n
        out[f"logits_{name}"] = np.stack(logits)
        R.close()
    np.savez_compressed(os.path.join(HERE, "tiny_logits.npz"), **out)

    # block codec known answers on the ggml test signal
    L = ref.ref_lib()
    n = 4096
    sig = (0.1 + 2.0 * np.cos(np.arange(n, dtype=np.float32))).astype(np.float32)
    sig2 = (0.1 + 2.0 * np.cos(np.arange(n, dtype=np.float32) + 1.0)).astype(np.float32)
    qout["signal_n"] = np.array([n])
    for name, t in (("q4_0", 2), ("q4_1", 3), ("q5_0", 6), ("q5_1", 7), ("q8_0", 8), ("q8_1", 9)):
        nbytes = n // 32 * L.ref_type_size(t)
        buf = np.zeros(nbytes, dtype=np.uint8)
        L.ref_from_float_reference(t, sig, buf, n)
        qout[f"codec_ref_{name}"] = buf.copy()
        buf2 = np.zeros(nbytes, dtype=np.uint8)
        L.ref_from_float(t, sig, buf2, n)           # the SIMD quantiser (what mul_mat applies to src1)
        qout[f"codec_simd_{name}"] = buf2.copy()
        if t != 9:
            deq = np.zeros(n, dtype=np.float32)
            L.ref_to_float(t, buf, deq, n)
            qout[f"dequant_{name}"] = deq
            vt = L.ref_vec_dot_type(t)
            ab = np.zeros(n // 32 * L.ref_type_size(vt), dtype=np.uint8)
            L.ref_from_float(vt, sig2, ab, n)
            qout[f"vecdot_{name}"] = np.array([L.ref_vec_dot(t, n, buf, ab)], dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, "quantize_ref.npz"), **qout)

    # lookup tables, probed through the reference ops on every fp16 input
    allh = np.arange(65536, dtype=np.uint16).view(np.float16).astype(np.float32)
    g = np.zeros(65536, dtype=np.float32)
    finite = np.isfinite(allh)
    x = np.where(finite, allh, 0).astype(np.float32)
    L.ref_gelu(x, g, 65536)
    np.savez_compressed(os.path.join(HERE, "tables.npz"), gelu_in=x, gelu_out=g)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()

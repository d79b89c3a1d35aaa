"""CPU tier: convert.py (HuggingFace checkpoint -> ggml .bin) against the reference's own converter
(/root/reference/convert.py) on a small checkpoint with the real tensor names -- byte for byte -- and against committed
hashes of the reference's output where /root/reference does not exist (tests/golden/convert_golden.json, written by
`python tests/test_convert.py --regen`)."""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/convert.py"
GOLD = os.path.join(ROOT, "tests", "golden", "convert_golden.json")


def make_checkpoint(d):
    """a 2-layer checkpoint with BioGPT's tensor names / shapes (HuggingFace order), a vocabulary and a merges file"""
    import torch
    from conftest import gf
    hp = gf.HParams(n_vocab=96, n_layer=2, n_head=4, n_positions=32, d_ff=64, d_model=32)
    rng = np.random.default_rng(7)
    sd = {}
    for name, shape, _ in gf.tensor_manifest(hp):
        a = (rng.standard_normal(shape) * 0.05).astype(np.float32)
        if name == "biogpt.layer_norm.bias":
            a = a.reshape(1, -1)                         # a squeezable dimension: the converter squeezes
        sd[name] = torch.from_numpy(a)
    os.makedirs(d, exist_ok=True)
    torch.save(sd, os.path.join(d, "pytorch_model.bin"))
    json.dump({"vocab_size": hp.n_vocab, "num_hidden_layers": hp.n_layer, "num_attention_heads": hp.n_head,
               "max_position_embeddings": hp.n_positions, "intermediate_size": hp.d_ff, "hidden_size": hp.d_model},
              open(os.path.join(d, "config.json"), "w"))
    words = ["<s>", "<pad>", "</s>", "<unk>"] + [f"w{i}</w>" for i in range(hp.n_vocab - 5)] + ["été</w>"]
    vocab = {w: i for i, w in enumerate(words)}
    json.dump(dict(reversed(list(vocab.items()))), open(os.path.join(d, "vocab.json"), "w", encoding="utf-8"), ensure_ascii=False)
    with open(os.path.join(d, "merges.txt"), "w", encoding="utf-8") as f:
        f.write("#version: 0.2\n")
        for i in range(40):
            f.write(f"a{i} b{i} {i}\n")
    return hp


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def run_ours(src, out, extra):
    subprocess.run([sys.executable, os.path.join(ROOT, "convert.py"), "--dir-model", src, "--out-dir", out] + extra,
                   check=True, capture_output=True)
    return os.path.join(out, "ggml-model.bin")


@pytest.mark.parametrize("f16", [False, True])
def test_convert_equals_reference_converter(tmp_path, f16):
    src = str(tmp_path / "hf")
    make_checkpoint(src)
    extra = ["--use-f16"] if f16 else []
    ours = run_ours(src, str(tmp_path / "ours"), extra)
    if os.path.exists(REF):
        subprocess.run([sys.executable, REF, "--dir-model", src, "--out-dir", str(tmp_path / "ref")] + extra, check=True, capture_output=True)
        a, b = open(ours, "rb").read(), open(str(tmp_path / "ref" / "ggml-model.bin"), "rb").read()
        assert len(a) == len(b) and a == b, f"first difference at byte {next(i for i, (x, y) in enumerate(zip(a, b)) if x != y)}"
    gold = json.load(open(GOLD))
    assert sha(ours) == gold["f16" if f16 else "f32"], "converter output changed against the committed hash of the reference's output"


def test_converted_file_parses_and_quantises_directly(tmp_path):
    from conftest import gf
    src = str(tmp_path / "hf")
    hp = make_checkpoint(src)
    f32 = gf.read_model(run_ours(src, str(tmp_path / "o32"), []))
    assert (f32.hparams.n_vocab, f32.hparams.n_layer, f32.hparams.d_model, f32.hparams.ftype) == (hp.n_vocab, hp.n_layer, hp.d_model, 0)
    q = gf.read_model(run_ours(src, str(tmp_path / "oq"), ["--ftype", "q5_1"]))
    assert q.hparams.ftype == 9
    w = f32.f32("biogpt.layers.1.fc1.weight")
    assert np.array_equal(q.raw("biogpt.layers.1.fc1.weight"), np.frombuffer(gf.encode_tensor(w, gf.GGML_TYPE_Q5_1), np.uint8))
    assert np.array_equal(q.f32("biogpt.layers.1.fc1.bias"), f32.f32("biogpt.layers.1.fc1.bias"))       # 1-D tensors stay f32


if __name__ == "__main__" and "--regen" in sys.argv:
    import tempfile
    assert os.path.exists(REF)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    out = {}
    with tempfile.TemporaryDirectory() as t:
        make_checkpoint(os.path.join(t, "hf"))
        for key, extra in (("f32", []), ("f16", ["--use-f16"])):
            subprocess.run([sys.executable, REF, "--dir-model", os.path.join(t, "hf"), "--out-dir", os.path.join(t, key)] + extra, check=True, capture_output=True)
            out[key] = sha(os.path.join(t, key, "ggml-model.bin"))
    json.dump(out, open(GOLD, "w"), indent=1)
    print(out)

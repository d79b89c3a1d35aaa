#!/usr/bin/env python
"""convert.py -- HuggingFace BioGPT checkpoint -> the ggml `.bin` the loader reads (the reference's offline step,
/root/reference/convert.py:1-119; file layout: biogpt.cpp:27-453).

    python convert.py --dir-model <dir with config.json, vocab.json, merges.txt, pytorch_model.bin> --out-dir <dir> [--use-f16]

writes <out-dir>/ggml-model.bin, byte-identical to the reference's converter on the same checkpoint
(tests/test_convert.py runs both).  As the reference does: header = magic + 7 hparams (ftype = 0 / 1), the vocabulary in id
order, the merges file split on newlines with the last piece dropped (a "#version" line stays a merge), then every tensor of
the checkpoint in checkpoint order, squeezed; with --use-f16 the 2-D tensors whose name ends in ".weight" become fp16,
everything else f32.  Extension: --ftype q4_0|q4_1|q5_0|q5_1|q8_0 quantises the 2-D weights directly (what the reference's
`quantize` tool does to the f32 file, biogpt.cpp:459-621; same blocks, tests/test_host_lib.py pins the codecs)."""
from __future__ import annotations

import argparse
import json
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from _bootstrap import load_pkg  # noqa: E402

GGML_MAGIC = 0x67676D6C
FTYPES = {"f32": 0, "f16": 1, "q4_0": 2, "q4_1": 3, "q8_0": 7, "q5_0": 8, "q5_1": 9}


def load_checkpoint(dir_model: str):
    """name -> numpy array, in checkpoint order"""
    pt = os.path.join(dir_model, "pytorch_model.bin")
    st = os.path.join(dir_model, "model.safetensors")
    if os.path.exists(pt):
        import torch
        ck = torch.load(pt, map_location="cpu")
        return [(k, v.squeeze().numpy()) for k, v in ck.items()]
    if os.path.exists(st):
        from safetensors.numpy import load_file
        return [(k, np.squeeze(v)) for k, v in load_file(st).items()]
    raise FileNotFoundError(f"no pytorch_model.bin / model.safetensors in {dir_model}")


def write_header(f, cfg: dict, ftype: int) -> None:
    f.write(struct.pack("<i", GGML_MAGIC))
    f.write(struct.pack("<7i", cfg["vocab_size"], cfg["num_hidden_layers"], cfg["num_attention_heads"],
                        cfg["max_position_embeddings"], cfg["intermediate_size"], cfg["hidden_size"], ftype))


def write_strings(f, items) -> None:
    f.write(struct.pack("<i", len(items)))
    for s in items:
        b = s.encode("utf-8")
        f.write(struct.pack("<i", len(b)))
        f.write(b)


def convert(dir_model: str, out_path: str, ftype_name: str = "f32", verbose: bool = True) -> None:
    gf = load_pkg().ggml_file
    ftype = FTYPES[ftype_name]
    wtype = gf.FTYPE_TO_TYPE[ftype]
    cfg = json.load(open(os.path.join(dir_model, "config.json"), encoding="utf-8"))
    vocab = json.load(open(os.path.join(dir_model, "vocab.json"), encoding="utf-8"))
    merges_raw = open(os.path.join(dir_model, "merges.txt"), encoding="utf-8").read().split("\n")[:-1]
    with open(out_path, "wb") as f:
        write_header(f, cfg, ftype)
        write_strings(f, [tok for tok, _ in sorted(vocab.items(), key=lambda kv: kv[1])])
        write_strings(f, [" ".join(line.split()[:2]) for line in merges_raw])
        for name, a in load_checkpoint(dir_model):
            is_mat = a.ndim == 2 and (name.endswith(".weight") if ftype == 1 else "weight" in name)
            ttype = wtype if (is_mat and ftype != 0) else gf.GGML_TYPE_F32
            if verbose:
                print(f"{name}: {tuple(a.shape)} -> ggml type {ttype}")
            nb = name.encode("utf-8")
            f.write(struct.pack("<3i", a.ndim, len(nb), ttype))
            for i in range(a.ndim):
                f.write(struct.pack("<i", a.shape[a.ndim - 1 - i]))
            f.write(nb)
            f.write(gf.encode_tensor(np.ascontiguousarray(a, dtype=np.float32), ttype))


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--dir-model", required=True)
    ap.add_argument("--out-dir", required=True)
    ap.add_argument("--use-f16", action="store_true")
    ap.add_argument("--ftype", choices=sorted(FTYPES), default=None, help="extension: write this format directly")
    a = ap.parse_args()
    os.makedirs(a.out_dir, exist_ok=True)
    ft = a.ftype or ("f16" if a.use_f16 else "f32")
    convert(a.dir_model, os.path.join(a.out_dir, "ggml-model.bin"), ft)
    print("Done.")


if __name__ == "__main__":
    main()

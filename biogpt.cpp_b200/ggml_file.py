"""ggml `.bin` model files for BioGPT: layout constants, reader, writer, block codecs.

The on-disk layout is the reference's contract (written by /root/reference/convert.py:53-98,
parsed by /root/reference/biogpt.cpp:41-434, rewritten by the quantize tool
biogpt.cpp:459-621).  Little-endian throughout:

    u32  magic 0x67676d6c
    i32  n_vocab, n_layer, n_head, n_positions, d_ff, d_model, ftype
    i32  n_vocab ; n_vocab x { u32 len ; bytes }
    i32  n_merges (must be 40000, biogpt.h:27) ; n_merges x { u32 len ; "a b" }
    per tensor: i32 n_dims, i32 name_len, i32 ggml_type, i32 ne[n_dims] (innermost first),
                name bytes, raw data

The numpy block quantisers restate `quantize_row_q*_reference` (ggml.c:892-1094), the
deterministic codecs the reference's quantize tool uses for weights.  This module is host
tooling (fixture generation, file inspection); the device never sees numpy.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

MAGIC = 0x67676D6C
N_MERGES = 40000  # biogpt.h:27 -- the loader rejects any other count (biogpt.cpp:120)

# ggml_type (ggml.h:306-314)
GGML_TYPE_F32, GGML_TYPE_F16 = 0, 1
GGML_TYPE_Q4_0, GGML_TYPE_Q4_1 = 2, 3
GGML_TYPE_Q5_0, GGML_TYPE_Q5_1 = 6, 7
GGML_TYPE_Q8_0, GGML_TYPE_Q8_1 = 8, 9

# ggml_ftype (ggml.h:337-345) -> ggml_type (ggml.c:4500-4523)
FTYPE_TO_TYPE = {0: GGML_TYPE_F32, 1: GGML_TYPE_F16, 2: GGML_TYPE_Q4_0, 3: GGML_TYPE_Q4_1,
                 7: GGML_TYPE_Q8_0, 8: GGML_TYPE_Q5_0, 9: GGML_TYPE_Q5_1}
TYPE_TO_FTYPE = {v: k for k, v in FTYPE_TO_TYPE.items()}
FTYPE_BY_NAME = {"f32": 0, "f16": 1, "q4_0": 2, "q4_1": 3, "q8_0": 7, "q5_0": 8, "q5_1": 9}
NAME_BY_FTYPE = {v: k for k, v in FTYPE_BY_NAME.items()}

QK = 32
# bytes per block of `blck` elements (ggml.c:844-889)
TYPE_SIZE = {GGML_TYPE_F32: 4, GGML_TYPE_F16: 2, GGML_TYPE_Q4_0: 18, GGML_TYPE_Q4_1: 20,
             GGML_TYPE_Q5_0: 22, GGML_TYPE_Q5_1: 24, GGML_TYPE_Q8_0: 34, GGML_TYPE_Q8_1: 40}
BLCK_SIZE = {GGML_TYPE_F32: 1, GGML_TYPE_F16: 1, GGML_TYPE_Q4_0: 32, GGML_TYPE_Q4_1: 32,
             GGML_TYPE_Q5_0: 32, GGML_TYPE_Q5_1: 32, GGML_TYPE_Q8_0: 32, GGML_TYPE_Q8_1: 32}


def row_bytes(ggml_type: int, k: int) -> int:
    return k // BLCK_SIZE[ggml_type] * TYPE_SIZE[ggml_type]


@dataclass
class HParams:
    n_vocab: int = 42384
    n_layer: int = 24
    n_head: int = 16
    n_positions: int = 1024
    d_ff: int = 4096
    d_model: int = 1024
    ftype: int = 0

    def pack(self) -> bytes:
        return struct.pack("<7i", self.n_vocab, self.n_layer, self.n_head, self.n_positions,
                           self.d_ff, self.d_model, self.ftype)


BASE = HParams()
TINY = HParams(n_vocab=256, n_layer=2, n_head=4, n_positions=64, d_ff=128, d_model=64)
SMALL = HParams(n_vocab=1000, n_layer=3, n_head=4, n_positions=256, d_ff=1024, d_model=256)
# BioGPT-base layer shapes (what the generation-4 persistent kernel is specialised for) with few
# layers and a small vocabulary, so the oracle walks the whole 1024-position context in seconds
NARROW = HParams(n_vocab=3001, n_layer=2, n_head=16, n_positions=1024, d_ff=4096, d_model=1024)


# --------------------------------------------------------------------------------------
# tensor manifest (names are the loader's lookup keys, biogpt.cpp:258-317)
# --------------------------------------------------------------------------------------

def tensor_manifest(hp: HParams) -> List[Tuple[str, Tuple[int, ...], bool]]:
    """[(name, torch-style shape (out, in) or (n,), is_weight_matrix)] in convert.py order."""
    d, ff, v = hp.d_model, hp.d_ff, hp.n_vocab
    out: List[Tuple[str, Tuple[int, ...], bool]] = [
        ("biogpt.embed_tokens.weight", (v, d), True),
        ("biogpt.embed_positions.weight", (d + 2, d), True),  # d_model+2 rows, biogpt.cpp:264
    ]
    for i in range(hp.n_layer):
        p = f"biogpt.layers.{i}."
        for proj in ("k_proj", "v_proj", "q_proj", "out_proj"):
            out.append((p + f"self_attn.{proj}.weight", (d, d), True))
            out.append((p + f"self_attn.{proj}.bias", (d,), False))
        out.append((p + "self_attn_layer_norm.weight", (d,), False))
        out.append((p + "self_attn_layer_norm.bias", (d,), False))
        out.append((p + "fc1.weight", (ff, d), True))
        out.append((p + "fc1.bias", (ff,), False))
        out.append((p + "fc2.weight", (d, ff), True))
        out.append((p + "fc2.bias", (d,), False))
        out.append((p + "final_layer_norm.weight", (d,), False))
        out.append((p + "final_layer_norm.bias", (d,), False))
    out.append(("biogpt.layer_norm.weight", (d,), False))
    out.append(("biogpt.layer_norm.bias", (d,), False))
    out.append(("output_projection.weight", (v, d), True))
    return out


def synth_tensors(hp: HParams, seed: int = 1234, w_std: float = 0.02) -> Dict[str, np.ndarray]:
    """Synthetic f32 weights: N(0, w_std^2) matrices, LN weight 1+0.05 N(0,1),
    biases 0.02 N(0,1)."""
    rng = np.random.default_rng(seed)
    t: Dict[str, np.ndarray] = {}
    for name, shape, is_mat in tensor_manifest(hp):
        if is_mat:
            a = rng.standard_normal(shape, dtype=np.float32) * np.float32(w_std)
        elif name.endswith("layer_norm.weight"):
            a = np.float32(1.0) + np.float32(0.05) * rng.standard_normal(shape, dtype=np.float32)
        else:
            a = np.float32(0.02) * rng.standard_normal(shape, dtype=np.float32)
        t[name] = np.ascontiguousarray(a, dtype=np.float32)
    return t


def synth_vocab(n_vocab: int) -> List[bytes]:
    """SURVEY appendix A recipe: specials, a..z, then tok{i}; ids `2 4 5 6` == "a b c"."""
    words = ["<s>", "<pad>", "</s>", "<unk>"] + [chr(ord("a") + i) + "</w>" for i in range(26)]
    i = 0
    while len(words) < n_vocab:
        words.append(f"tok{i}</w>")
        i += 1
    return [w.encode() for w in words[:n_vocab]]


def synth_merges() -> List[bytes]:
    return [f"zz{i} yy{i}".encode() for i in range(N_MERGES)]


# --------------------------------------------------------------------------------------
# block codecs (numpy restatements of ggml.c:892-1094 and 1536-1646)
# --------------------------------------------------------------------------------------

def _f32(x):
    return np.asarray(x, dtype=np.float32)


def _fma32(a, b, c):
    """float32 fused multiply-add, emulated exactly: a*b is exact in float64 (48-bit
    product), the float64 sum is then rounded once more to float32.  The double rounding can
    only differ from a true fma when the float64 sum lands exactly on a float32 tie, which
    needs >29 cancelling bits; the oracle tests pin this against the reference build."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def _absmax_signed(xb: np.ndarray) -> np.ndarray:
    """value with the largest |v| in each block, first occurrence wins (strict `<`)."""
    idx = np.argmax(np.abs(xb), axis=1)
    return xb[np.arange(xb.shape[0]), idx]


def _trunc_i8(v: np.ndarray) -> np.ndarray:
    # (int8_t)(float): C truncation toward zero; values here are always within int8 range
    return np.trunc(v).astype(np.int32)


def quantize_q4_0(x: np.ndarray, fused: bool = True) -> np.ndarray:
    """`fused`: gcc contracts `x*id + 8.5f` to an fma in the reference build
    (-O3 -mfma, GNU C default -ffp-contract=fast); pinned in tests/test_codecs.py."""
    xb = _f32(x).reshape(-1, QK)
    mx = _absmax_signed(xb)
    d = (mx / np.float32(-8)).astype(np.float32)
    with np.errstate(divide="ignore"):
        idv = np.where(d != 0, np.float32(1.0) / d, np.float32(0)).astype(np.float32)
    if fused:
        v = _fma32(xb, idv[:, None], np.float32(8.5))
    else:
        v = (xb * idv[:, None]).astype(np.float32) + np.float32(8.5)
    q = np.minimum(15, _trunc_i8(v)).astype(np.uint8)
    out = np.zeros((xb.shape[0], 18), dtype=np.uint8)
    out[:, 0:2] = d.astype(np.float16).view(np.uint8).reshape(-1, 2)
    out[:, 2:] = (q[:, :16] & 0x0F) | ((q[:, 16:] & 0x0F) << 4)
    return out.reshape(-1)


def quantize_q4_1(x: np.ndarray, fused: bool = True) -> np.ndarray:
    xb = _f32(x).reshape(-1, QK)
    mn, mx = xb.min(axis=1), xb.max(axis=1)
    d = ((mx - mn) / np.float32(15)).astype(np.float32)
    with np.errstate(divide="ignore"):
        idv = np.where(d != 0, np.float32(1.0) / d, np.float32(0)).astype(np.float32)
    t = (xb - mn[:, None]).astype(np.float32)
    if fused:
        v = _fma32(t, idv[:, None], np.float32(0.5))
    else:
        v = (t * idv[:, None]).astype(np.float32) + np.float32(0.5)
    q = np.minimum(15, _trunc_i8(v)).astype(np.uint8)
    out = np.zeros((xb.shape[0], 20), dtype=np.uint8)
    out[:, 0:2] = d.astype(np.float16).view(np.uint8).reshape(-1, 2)
    out[:, 2:4] = mn.astype(np.float16).view(np.uint8).reshape(-1, 2)
    out[:, 4:] = (q[:, :16] & 0x0F) | ((q[:, 16:] & 0x0F) << 4)
    return out.reshape(-1)


def _pack_qh(q: np.ndarray) -> np.ndarray:
    bits = ((q >> 4) & 1).astype(np.uint32)  # element j -> bit j (j<16 low half, j>=16 high)
    sh = np.arange(32, dtype=np.uint32)
    return (bits << sh).sum(axis=1, dtype=np.uint64).astype(np.uint32)


def quantize_q5_0(x: np.ndarray, fused: bool = True) -> np.ndarray:
    xb = _f32(x).reshape(-1, QK)
    mx = _absmax_signed(xb)
    d = (mx / np.float32(-16)).astype(np.float32)
    with np.errstate(divide="ignore"):
        idv = np.where(d != 0, np.float32(1.0) / d, np.float32(0)).astype(np.float32)
    if fused:
        v = _fma32(xb, idv[:, None], np.float32(16.5))
    else:
        v = (xb * idv[:, None]).astype(np.float32) + np.float32(16.5)
    q = np.minimum(31, _trunc_i8(v)).astype(np.uint8)
    out = np.zeros((xb.shape[0], 22), dtype=np.uint8)
    out[:, 0:2] = d.astype(np.float16).view(np.uint8).reshape(-1, 2)
    out[:, 2:6] = _pack_qh(q).view(np.uint8).reshape(-1, 4)
    out[:, 6:] = (q[:, :16] & 0x0F) | ((q[:, 16:] & 0x0F) << 4)
    return out.reshape(-1)


def quantize_q5_1(x: np.ndarray, fused: bool = True) -> np.ndarray:
    xb = _f32(x).reshape(-1, QK)
    mn, mx = xb.min(axis=1), xb.max(axis=1)
    d = ((mx - mn) / np.float32(31)).astype(np.float32)
    with np.errstate(divide="ignore"):
        idv = np.where(d != 0, np.float32(1.0) / d, np.float32(0)).astype(np.float32)
    t = (xb - mn[:, None]).astype(np.float32)
    if fused:
        v = _fma32(t, idv[:, None], np.float32(0.5))
    else:
        v = (t * idv[:, None]).astype(np.float32) + np.float32(0.5)
    q = np.trunc(v).astype(np.int32).astype(np.uint8)  # (uint8_t)(x0 + 0.5f), no clamp
    out = np.zeros((xb.shape[0], 24), dtype=np.uint8)
    out[:, 0:2] = d.astype(np.float16).view(np.uint8).reshape(-1, 2)
    out[:, 2:4] = mn.astype(np.float16).view(np.uint8).reshape(-1, 2)
    out[:, 4:8] = _pack_qh(q).view(np.uint8).reshape(-1, 4)
    out[:, 8:] = (q[:, :16] & 0x0F) | ((q[:, 16:] & 0x0F) << 4)
    return out.reshape(-1)


def _roundf(v: np.ndarray) -> np.ndarray:
    # C roundf: half away from zero
    return np.where(v >= 0, np.floor(v + np.float32(0.5)), np.ceil(v - np.float32(0.5))).astype(np.float32)


def quantize_q8_0(x: np.ndarray, fused: bool = True) -> np.ndarray:
    xb = _f32(x).reshape(-1, QK)
    amax = np.abs(xb).max(axis=1)
    d = (amax / np.float32(127)).astype(np.float32)
    with np.errstate(divide="ignore"):
        idv = np.where(d != 0, np.float32(1.0) / d, np.float32(0)).astype(np.float32)
    v = (xb * idv[:, None]).astype(np.float32)
    # roundf on float32: |v| <= 127.x so v +- 0.5 is exact enough; do it in float64 to be safe
    v64 = v.astype(np.float64)
    q = np.where(v64 >= 0, np.floor(v64 + 0.5), np.ceil(v64 - 0.5)).astype(np.int8)
    out = np.zeros((xb.shape[0], 34), dtype=np.uint8)
    out[:, 0:2] = d.astype(np.float16).view(np.uint8).reshape(-1, 2)
    out[:, 2:] = q.view(np.uint8)
    return out.reshape(-1)


QUANTIZERS = {GGML_TYPE_Q4_0: quantize_q4_0, GGML_TYPE_Q4_1: quantize_q4_1,
              GGML_TYPE_Q5_0: quantize_q5_0, GGML_TYPE_Q5_1: quantize_q5_1,
              GGML_TYPE_Q8_0: quantize_q8_0}


def _h2f(b: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(b).view(np.float16).astype(np.float32).reshape(-1)


def dequantize(ggml_type: int, raw: np.ndarray, n: int) -> np.ndarray:
    """`dequantize_row_q*` (ggml.c:1536-1646), unfused float32 arithmetic."""
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    if ggml_type == GGML_TYPE_F32:
        return raw.view(np.float32)[:n].copy()
    if ggml_type == GGML_TYPE_F16:
        return raw.view(np.float16)[:n].astype(np.float32)
    nb = n // QK
    blk = raw.reshape(nb, TYPE_SIZE[ggml_type])
    d = _h2f(blk[:, 0:2])[:, None]
    if ggml_type == GGML_TYPE_Q8_0:
        q = blk[:, 2:].view(np.int8).astype(np.float32)
        return (q * d).astype(np.float32).reshape(-1)
    if ggml_type in (GGML_TYPE_Q4_0, GGML_TYPE_Q5_0):
        off = 2
        m = None
    else:
        off = 4
        m = _h2f(blk[:, 2:4])[:, None]
    if ggml_type in (GGML_TYPE_Q5_0, GGML_TYPE_Q5_1):
        qh = np.ascontiguousarray(blk[:, off:off + 4]).view(np.uint32).reshape(nb, 1)
        hb = ((qh >> np.arange(32, dtype=np.uint32)[None, :]) & 1).astype(np.int32) << 4
        off += 4
    else:
        hb = np.zeros((nb, 32), dtype=np.int32)
    qs = blk[:, off:off + 16]
    q = np.concatenate([qs & 0x0F, qs >> 4], axis=1).astype(np.int32) | hb
    if ggml_type == GGML_TYPE_Q4_0:
        q = q - 8
    if ggml_type == GGML_TYPE_Q5_0:
        q = q - 16
    y = (q.astype(np.float32) * d).astype(np.float32)
    if m is not None:
        y = (y + m).astype(np.float32)
    return y.reshape(-1)


def encode_tensor(a: np.ndarray, ggml_type: int) -> bytes:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if ggml_type == GGML_TYPE_F32:
        return a.tobytes()
    if ggml_type == GGML_TYPE_F16:
        return a.astype(np.float16).tobytes()
    return QUANTIZERS[ggml_type](a.reshape(-1)).tobytes()


# --------------------------------------------------------------------------------------
# writer / reader
# --------------------------------------------------------------------------------------

def write_model(path: str, hp: HParams, tensors: Dict[str, np.ndarray], ftype: int,
                vocab: Optional[List[bytes]] = None, merges: Optional[List[bytes]] = None) -> None:
    """Write a `.bin` the reference loader accepts.  ftype 0: all f32.  ftype 1: 2-D `.weight`
    tensors f16 (convert.py:62-66).  Quantised ftypes: 2-D tensors whose name contains "weight"
    are block-quantised, 1-D stay f32 (biogpt.cpp:523)."""
    wtype = FTYPE_TO_TYPE[ftype]
    vocab = vocab if vocab is not None else synth_vocab(hp.n_vocab)
    merges = merges if merges is not None else synth_merges()
    assert len(vocab) == hp.n_vocab and len(merges) == N_MERGES
    hp = HParams(**{**hp.__dict__, "ftype": ftype})
    with open(path, "wb") as f:
        f.write(struct.pack("<I", MAGIC))
        f.write(hp.pack())
        f.write(struct.pack("<i", len(vocab)))
        for w in vocab:
            f.write(struct.pack("<I", len(w)))
            f.write(w)
        f.write(struct.pack("<i", len(merges)))
        for m in merges:
            f.write(struct.pack("<I", len(m)))
            f.write(m)
        for name, shape, is_mat in tensor_manifest(hp):
            a = tensors[name]
            assert tuple(a.shape) == tuple(shape), (name, a.shape, shape)
            ttype = wtype if is_mat else GGML_TYPE_F32
            nb = name.encode()
            f.write(struct.pack("<3i", len(shape), len(nb), ttype))
            for i in range(len(shape)):
                f.write(struct.pack("<i", shape[len(shape) - 1 - i]))
            f.write(nb)
            f.write(encode_tensor(a, ttype))


@dataclass
class TensorEntry:
    name: str
    ggml_type: int
    ne: Tuple[int, ...]      # innermost first, as stored
    offset: int              # byte offset of the raw data in the file
    nbytes: int


@dataclass
class ModelFile:
    hparams: HParams
    vocab: List[bytes]
    merges: List[bytes]
    tensors: Dict[str, TensorEntry] = field(default_factory=dict)
    path: str = ""

    def raw(self, name: str) -> np.ndarray:
        e = self.tensors[name]
        return np.fromfile(self.path, dtype=np.uint8, count=e.nbytes, offset=e.offset)

    def f32(self, name: str) -> np.ndarray:
        e = self.tensors[name]
        n = int(np.prod(e.ne))
        return dequantize(e.ggml_type, self.raw(name), n).reshape(tuple(reversed(e.ne)))


def read_model(path: str) -> ModelFile:
    with open(path, "rb") as f:
        (magic,) = struct.unpack("<I", f.read(4))
        if magic != MAGIC:
            raise ValueError(f"{path}: bad magic {magic:#x}")
        hp = HParams(*struct.unpack("<7i", f.read(28)))
        (nv,) = struct.unpack("<i", f.read(4))
        vocab = []
        for _ in range(nv):
            (ln,) = struct.unpack("<I", f.read(4))
            vocab.append(f.read(ln))
        (nm,) = struct.unpack("<i", f.read(4))
        merges = []
        for _ in range(nm):
            (ln,) = struct.unpack("<I", f.read(4))
            merges.append(f.read(ln))
        mf = ModelFile(hp, vocab, merges, path=path)
        while True:
            hdr = f.read(12)
            if len(hdr) < 12:
                break
            n_dims, name_len, ttype = struct.unpack("<3i", hdr)
            ne = struct.unpack(f"<{n_dims}i", f.read(4 * n_dims))
            name = f.read(name_len).decode()
            n = int(np.prod(ne))
            nbytes = n // BLCK_SIZE[ttype] * TYPE_SIZE[ttype]
            mf.tensors[name] = TensorEntry(name, ttype, tuple(ne), f.tell(), nbytes)
            f.seek(nbytes, 1)
    return mf


def synth_tokens(n: int, n_vocab: int, seed: int = 0) -> np.ndarray:
    """first id 2 (`</s>`), then uniform ids in [4, n_vocab)."""
    rng = np.random.default_rng(seed)
    t = rng.integers(4, n_vocab, size=n, dtype=np.int32)
    t[0] = 2
    return t

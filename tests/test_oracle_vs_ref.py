"""CPU tier, builder container only: the C restatement against the UNMODIFIED reference compiled
in place from /root/reference (oracle/_ref/libbiogpt_ref.so).  Skipped where the reference build
is absent (the GPU box); there tests/test_oracle_golden.py carries the pin."""
import os

import numpy as np
import pytest

from conftest import FTYPES, ROOT, gf

pytestmark = pytest.mark.skipif(
    not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libbiogpt_ref.so")) and not os.path.isdir("/root/reference"),
    reason="reference build (oracle/_ref) not available on this machine")

TYPES = {"f32": 0, "f16": 1, "q4_0": 2, "q4_1": 3, "q5_0": 6, "q5_1": 7, "q8_0": 8}


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("ftype", FTYPES)
def test_eval_bit_exact_tiny(checkers, zoo, ftype):
    p = zoo.path("tiny", ftype)
    R, O = checkers.Ref(p), checkers.Oracle(p)
    toks = gf.synth_tokens(40, gf.TINY.n_vocab, seed=3)
    pos = 0
    for n in (4, 8, 1, 1, 3, 1, 7, 1, 1, 1):
        a, b = R.eval(toks[pos:pos + n], pos), O.eval(toks[pos:pos + n], pos)
        assert np.array_equal(_bits(a), _bits(b)), (ftype, pos, n)
        pos += n
    R.close(); O.close()


@pytest.mark.parametrize("ftype", ["f16", "q4_0", "q5_1"])
def test_eval_bit_exact_small(checkers, zoo, ftype):
    """d_model 256 / d_kv 64 / d_ff 1024: the vectorised (non-tail) dot paths"""
    p = zoo.path("small", ftype)
    R, O = checkers.Ref(p, n_batch=40), checkers.Oracle(p)   # measure pass must cover the 33-token batch
    toks = gf.synth_tokens(80, gf.SMALL.n_vocab, seed=5)
    pos = 0
    for n in (8, 8, 8, 8, 8, 1, 1, 1, 33, 1):
        a, b = R.eval(toks[pos:pos + n], pos), O.eval(toks[pos:pos + n], pos)
        assert np.array_equal(_bits(a), _bits(b)), (ftype, pos, n)
        pos += n
    R.close(); O.close()


@pytest.mark.parametrize("name", list(TYPES))
def test_mul_mat_bit_exact(checkers, name):
    t = TYPES[name]
    rng = np.random.default_rng(11)
    k, rows, n = 256, 40, 3
    w = (rng.standard_normal((rows, k)) * 0.05).astype(np.float32)
    x = rng.standard_normal((n, k)).astype(np.float32)
    wb = np.frombuffer(gf.encode_tensor(w, t), dtype=np.uint8).copy()
    y_ref = np.zeros((n, rows), dtype=np.float32)
    y_or = np.zeros((n, rows), dtype=np.float32)
    checkers.ref_lib().ref_mul_mat(t, wb, x, y_ref, k, rows, n, 4)
    checkers.oracle_lib().bo_mul_mat(t, wb, x, y_or, k, rows, n)
    assert np.array_equal(_bits(y_ref), _bits(y_or))


def test_norm_softmax_gelu_bit_exact(checkers):
    rng = np.random.default_rng(2)
    R, O = checkers.ref_lib(), checkers.oracle_lib()
    for nc in (64, 1024, 4096):
        x = (rng.standard_normal(nc) * 3).astype(np.float32)
        a, b = np.zeros(nc, np.float32), np.zeros(nc, np.float32)
        R.ref_norm(x, a, nc, 1, 1e-5); O.bo_norm(x, b, nc, 1e-5)
        assert np.array_equal(_bits(a), _bits(b))
        R.ref_soft_max(x, a, nc, 1); O.bo_soft_max(x, b, nc)
        assert np.array_equal(_bits(a), _bits(b))
        R.ref_gelu(x, a, nc); O.bo_gelu(x, b, nc)
        assert np.array_equal(_bits(a), _bits(b))


def test_activation_quantisers_bit_exact(checkers):
    rng = np.random.default_rng(4)
    R, O = checkers.ref_lib(), checkers.oracle_lib()
    x = (rng.standard_normal(4096) * 2).astype(np.float32)
    x[64:96] = 0.0  # an all-zero block (id = 0 branch)
    a, b = np.zeros(4096 // 32 * 34, np.uint8), np.zeros(4096 // 32 * 34, np.uint8)
    R.ref_from_float(8, x, a, 4096); O.bo_quantize_row_q8_0(x, b, 4096)
    assert np.array_equal(a, b)
    a, b = np.zeros(4096 // 32 * 40, np.uint8), np.zeros(4096 // 32 * 40, np.uint8)
    R.ref_from_float(9, x, a, 4096); O.bo_quantize_row_q8_1(x, b, 4096)
    assert np.array_equal(a, b)

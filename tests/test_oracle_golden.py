"""CPU tier: the C restatement (oracle/biogpt_oracle.c) against the committed fixtures that were
generated from the UNMODIFIED reference (tests/golden/make_golden.py).  This is what pins the
oracle on machines where /root/reference does not exist (the GPU box)."""
import hashlib
import os

import numpy as np
import pytest

from conftest import FTYPES, ROOT, gf

GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "tiny_logits.npz"))


@pytest.mark.parametrize("ftype", FTYPES)
def test_model_file_bytes_match_fixture(zoo, gold, ftype):
    """the `.bin` every test rebuilds is byte-identical to the one the golden logits came from"""
    p = zoo.path("tiny", ftype)
    sha = hashlib.sha256(open(p, "rb").read()).digest()
    assert sha == gold[f"sha_{ftype}"].tobytes()


def test_numpy_quantizers_match_reference_tool(gold):
    """ggml_file's block quantisers == bytes written by the reference's quantize tool"""
    q = np.load(os.path.join(GOLD, "quantize_ref.npz"))
    for ftype in ("q4_0", "q4_1", "q5_0", "q5_1", "q8_0"):
        assert gold[f"sha_{ftype}"].tobytes() == q[f"sha_{ftype}"].tobytes(), ftype


@pytest.mark.parametrize("ftype", FTYPES)
def test_oracle_logits_equal_reference_logits(checkers, zoo, gold, ftype):
    """bit-for-bit: every eval of the schedule (prompt batches of 5 and 3, decode steps, a batch of 2)"""
    O = checkers.Oracle(zoo.path("tiny", ftype))
    toks, sched, want = gold["tokens"], gold["schedule"], gold[f"logits_{ftype}"]
    pos = 0
    for i, (n, n_past) in enumerate(sched):
        got = O.eval(toks[pos:pos + n], int(n_past))
        pos += n
        assert np.array_equal(got.view(np.uint32), want[i].view(np.uint32)), (ftype, i, np.abs(got - want[i]).max())
    O.close()


def test_codecs_on_ggml_test_signal(checkers):
    """block codecs on x_i = 0.1 + 2 cos(i) (ggml/tests/test-quantize-fns.cpp:26-30, 97):
    numpy weight quantisers == from_float_reference, oracle activation quantisers == the AVX
    from_float, dequantisers and vec_dot == the reference's, and the reference test's own
    thresholds (rmse < 0.002, dot error < 0.02) hold."""
    q = np.load(os.path.join(GOLD, "quantize_ref.npz"))
    n = int(q["signal_n"][0])
    sig = (0.1 + 2.0 * np.cos(np.arange(n, dtype=np.float32))).astype(np.float32)
    sig2 = (0.1 + 2.0 * np.cos(np.arange(n, dtype=np.float32) + 1.0)).astype(np.float32)
    L = checkers.oracle_lib()
    types = {"q4_0": 2, "q4_1": 3, "q5_0": 6, "q5_1": 7, "q8_0": 8}
    for name, t in types.items():
        mine = gf.QUANTIZERS[t](sig)
        assert np.array_equal(mine, q[f"codec_ref_{name}"]), name
        deq = np.zeros(n, dtype=np.float32)
        L.bo_dequantize_row(t, mine, deq, n)
        assert np.array_equal(deq, q[f"dequant_{name}"]), name
        assert np.array_equal(gf.dequantize(t, mine, n), q[f"dequant_{name}"]), name
        # array_rmse of the reference test is sqrt(sum of squares) / n (test-quantize-fns.cpp:33-40)
        rmse = np.sqrt(np.sum((deq.astype(np.float64) - sig) ** 2)) / n
        assert rmse < 0.002, (name, rmse)
        # activation side + dot
        vt_q81 = name in ("q4_1", "q5_1")
        act = np.zeros(n // 32 * (40 if vt_q81 else 34), dtype=np.uint8)
        (L.bo_quantize_row_q8_1 if vt_q81 else L.bo_quantize_row_q8_0)(sig2, act, n)
        dot = L.bo_vec_dot(t, n, mine, act)
        assert np.float32(dot) == q[f"vecdot_{name}"][0], name
        assert abs(dot - float(np.dot(sig.astype(np.float64), sig2))) / n < 0.02, name
    a80 = np.zeros(n // 32 * 34, dtype=np.uint8)
    L.bo_quantize_row_q8_0(sig, a80, n)
    assert np.array_equal(a80, q["codec_simd_q8_0"])
    a81 = np.zeros(n // 32 * 40, dtype=np.uint8)
    L.bo_quantize_row_q8_1(sig, a81, n)
    assert np.array_equal(a81, q["codec_simd_q8_1"])


def test_gelu_table_matches_reference(checkers, capi):
    """all 65536 entries, for both the oracle's and the product's (host libm) table builders"""
    z = np.load(os.path.join(GOLD, "tables.npz"))
    y = np.zeros(65536, dtype=np.float32)
    checkers.oracle_lib().bo_gelu(z["gelu_in"], y, 65536)
    assert np.array_equal(y.view(np.uint32), z["gelu_out"].view(np.uint32))
    og = np.zeros(65536, np.uint16); oe = np.zeros(65536, np.uint16)
    checkers.oracle_lib().bo_tables(og, oe)
    g, e = capi.build_tables()
    assert np.array_equal(g, og) and np.array_equal(e, oe)
    h = z["gelu_in"].astype(np.float16).view(np.uint16)
    assert np.array_equal(g[h].view(np.float16).astype(np.float32).view(np.uint32), z["gelu_out"].view(np.uint32))

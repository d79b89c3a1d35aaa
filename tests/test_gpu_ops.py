"""GPU tier: each device operator against the CPU oracle, through the C ABI, bit for bit.

The oracle restates the reference's AVX2 arithmetic order (oracle/biogpt_oracle.c); the kernels
follow the same order (bgpt_kernels.cuh, "lane order"), so every comparison here is exact
equality of the float bit patterns, not a tolerance."""
import numpy as np
import pytest

from conftest import gf

pytestmark = pytest.mark.gpu

TYPES = {"f32": 0, "f16": 1, "q4_0": 2, "q4_1": 3, "q5_0": 6, "q5_1": 7, "q8_0": 8}


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _report(name, got, want):
    bad = np.flatnonzero(_bits(got).ravel() != _bits(want).ravel())
    return f"{name}: {bad.size}/{got.size} differ, max|d|={np.abs(got - want).max():.3e}, first={bad[:5]}"


@pytest.mark.parametrize("name", list(TYPES))
@pytest.mark.parametrize("k", [64, 1024, 4096])
def test_quantize_act(checkers, capi, name, k):
    """mul_mat's src1 conversion: Q8_0 / Q8_1 blocks or fp16, incl. an all-zero block"""
    t = TYPES[name]
    rng = np.random.default_rng(k + t)
    x = (rng.standard_normal(k) * 2.5).astype(np.float32)
    x[32:64] = 0.0
    got = capi.op_quantize_act(t, x)
    O = checkers.oracle_lib()
    if t == 0:
        want = x.view(np.uint8)
    elif t == 1:
        h = np.zeros(k, np.uint16); O.bo_fp32_to_fp16_row(x, h, k); want = h.view(np.uint8)
    elif t in (2, 6, 8):
        want = np.zeros(k // 32 * 34, np.uint8); O.bo_quantize_row_q8_0(x, want, k)
    else:
        want = np.zeros(k // 32 * 40, np.uint8); O.bo_quantize_row_q8_1(x, want, k)
    assert np.array_equal(got, want), f"{name} k={k}: {np.flatnonzero(got != want)[:8]}"


@pytest.mark.parametrize("name", list(TYPES))
@pytest.mark.parametrize("shape", [(64, 24, 1), (1024, 1024, 1), (4096, 1024, 1), (1024, 4096, 3),
                                   (1024, 1000, 8), (4096, 264, 9), (1024, 2649 * 16, 1), (128, 8, 2)])
def test_mul_mat(checkers, capi, name, shape):
    """the seven weight formats x the BioGPT shapes (K 1024/4096, M 1024/4096/42384) and ragged
    ones (rows not a multiple of the tile, 9 token rows = one full tile + 1)"""
    k, rows, n = shape
    t = TYPES[name]
    rng = np.random.default_rng(k * 7 + rows + n + t)
    w = (rng.standard_normal((rows, k)) * 0.02).astype(np.float32)
    x = rng.standard_normal((n, k)).astype(np.float32)
    wb = np.frombuffer(gf.encode_tensor(w, t), dtype=np.uint8).copy()
    want = np.zeros((n, rows), dtype=np.float32)
    checkers.oracle_lib().bo_mul_mat(t, wb, x, want, k, rows, n)
    got = capi.op_mul_mat(t, wb, x, rows)
    assert np.array_equal(_bits(got), _bits(want)), _report(f"{name} {shape}", got, want)


def test_norm_rows_where_the_double_sums_round(checkers, capi):
    """LayerNorm's two sums run in double (ggml.c:11403-11420), sequentially in the reference and as a tree here.  A double sum of floats
    is EXACT -- hence independent of the order -- while every element is within 2^29 of the running sum's magnitude, which covers any
    activation row of the model; these rows leave that domain on purpose (20-40 binades of dynamic range, heavy cancellation, one
    dominant element), so the individual double additions do round and the two orders can differ in the last bits of the double.  The
    float results must still agree: the difference only shows if it moves the double across a float rounding boundary."""
    rng = np.random.default_rng(77)
    nc = 1024
    rows = []
    for e_lo in (-20, -30, -40):
        rows.append((rng.standard_normal(nc) * np.exp2(rng.uniform(e_lo, 4, nc))).astype(np.float32))
    big = (rng.standard_normal(nc) * 1e6).astype(np.float32)
    rows.append(np.concatenate([big[:nc // 2], -big[:nc // 2]]) + (rng.standard_normal(nc) * 1e-3).astype(np.float32))   # cancellation
    one = (rng.standard_normal(nc) * 1e-4).astype(np.float32); one[517] = 3e7; rows.append(one)                        # one dominant element
    rows.append((rng.standard_normal(nc) * np.exp2(rng.integers(-60, 10, nc).astype(np.float64))).astype(np.float32))
    x = np.stack(rows).astype(np.float32)
    want = np.zeros_like(x)
    for r in range(len(rows)):
        y = np.zeros(nc, np.float32)
        checkers.oracle_lib().bo_norm(x[r], y, nc, 1e-5)
        want[r] = y
    got = capi.op_norm(x, None, None, 1e-5)
    assert np.array_equal(_bits(got), _bits(want)), _report("norm, rounding double sums", got, want)


@pytest.mark.parametrize("nc", [64, 256, 1024, 4096])
def test_norm(checkers, capi, nc):
    rng = np.random.default_rng(nc)
    rows = 5
    x = (rng.standard_normal((rows, nc)) * 3 + 0.5).astype(np.float32)
    w = (1 + 0.05 * rng.standard_normal(nc)).astype(np.float32)
    b = (0.02 * rng.standard_normal(nc)).astype(np.float32)
    want = np.zeros_like(x)
    for r in range(rows):
        y = np.zeros(nc, np.float32)
        checkers.oracle_lib().bo_norm(x[r], y, nc, 1e-5)
        want[r] = y
    got = capi.op_norm(x, None, None, 1e-5)
    assert np.array_equal(_bits(got), _bits(want)), _report("norm", got, want)
    want_aff = ((w * want).astype(np.float32) + b).astype(np.float32)
    got_aff = capi.op_norm(x, w, b, 1e-5)
    assert np.array_equal(_bits(got_aff), _bits(want_aff)), _report("norm+affine", got_aff, want_aff)


def _oracle_attention(O, q, k, v, n_head):
    """softmax(K q) V per head, composed from the oracle's own ops (biogpt_oracle.c bo_eval)"""
    n, d = q.shape
    T = k.shape[0]
    dk = d // n_head
    out = np.zeros_like(q)
    sc = np.zeros(T, np.float32)
    for h in range(n_head):
        for i in range(n):
            qh = np.ascontiguousarray(q[i, h * dk:(h + 1) * dk])
            for t in range(T):
                sc[t] = O.bo_vec_dot_f32(dk, np.ascontiguousarray(k[t, h * dk:(h + 1) * dk]), qh)
            p = np.zeros(T, np.float32)
            O.bo_soft_max(sc, p, T)
            for c in range(dk):
                out[i, h * dk + c] = O.bo_vec_dot_f32(T, np.ascontiguousarray(v[:, h * dk + c]), p)
    return out


@pytest.mark.parametrize("cfg", [(64, 4, 1, 0), (64, 4, 3, 5), (256, 4, 1, 40), (256, 4, 2, 62), (1024, 16, 1, 97),
                                 (1024, 16, 1, 31), (256, 2, 1, 200), (128, 4, 1, 35),
                                 (256, 4, 8, 40), (256, 4, 5, 27), (1024, 16, 16, 17), (256, 4, 19, 77), (256, 4, 33, 0)])
def test_attention(checkers, capi, cfg):
    """un-masked attention: T = n_past + n below / at / above the 32-wide vector boundary, with a
    4-multiple tail and a <4 remainder; head dims 16, 32, 64, 128"""
    d, n_head, n, n_past = cfg
    rng = np.random.default_rng(d + n_past)
    T = n_past + n
    q = (rng.standard_normal((n, d)) * 0.5).astype(np.float32)
    k = rng.standard_normal((T, d)).astype(np.float32)
    v = rng.standard_normal((T, d)).astype(np.float32)
    O = checkers.oracle_lib()
    g, e = capi.build_tables()
    want = _oracle_attention(O, q, k, v, n_head)
    got = capi.op_attention(q, k, v, n_past, n_head, e)
    assert np.array_equal(_bits(got), _bits(want)), _report(f"attention {cfg}", got, want)


def test_gelu(checkers, capi):
    x = np.arange(65536, dtype=np.uint16).view(np.float16).astype(np.float32)
    x = np.where(np.isfinite(x), x, 0).astype(np.float32)
    x = np.concatenate([x, (np.random.default_rng(0).standard_normal(4096) * 4).astype(np.float32)])
    want = np.zeros_like(x)
    checkers.oracle_lib().bo_gelu(x, want, x.size)
    g, e = capi.build_tables()
    got = capi.op_gelu(x, g)
    assert np.array_equal(_bits(got), _bits(want))


@pytest.mark.parametrize("name", list(TYPES))
def test_dequantize_rows(checkers, capi, name):
    """get_rows on every storage type (embedding gather)"""
    t = TYPES[name]
    rng = np.random.default_rng(t)
    k, rows = 1024, 7
    w = (rng.standard_normal((rows, k)) * 0.02).astype(np.float32)
    wb = np.frombuffer(gf.encode_tensor(w, t), dtype=np.uint8).copy()
    want = np.zeros((rows, k), np.float32)
    rb = gf.row_bytes(t, k)
    for r in range(rows):
        y = np.zeros(k, np.float32)
        checkers.oracle_lib().bo_dequantize_row(t, wb[r * rb:(r + 1) * rb].copy(), y, k)
        want[r] = y
    got = capi.op_dequantize(t, wb, k, rows)
    assert np.array_equal(_bits(got), _bits(want))


@pytest.mark.parametrize("name", ["q4_0", "q4_1", "q5_0", "q5_1", "q8_0"])
@pytest.mark.parametrize("shape", [(1024, 256, 64), (4096, 136, 70), (1024, 1000, 200), (128, 128, 33), (64, 8, 5)])
def test_mul_mat_tensor_core(checkers, capi, name, shape):
    """tcgen05 prompt-batch matmul (csrc/bgpt_tc.cuh): integer block dots are exact, blocks are
    accumulated in f32 in block order -- so the result is the oracle's up to f32 summation order.
    Tolerance: 4 ulp-ish of the largest partial sum, |y_tc - y_oracle| <= 2e-6 * sum_b |term_b|,
    checked through the cheaper bound 3e-5 * max|y| + 1e-6."""
    k, rows, n = shape
    t = TYPES[name]
    rng = np.random.default_rng(k + rows * 3 + n + t)
    w = (rng.standard_normal((rows, k)) * 0.02).astype(np.float32)
    x = rng.standard_normal((n, k)).astype(np.float32)
    wb = np.frombuffer(gf.encode_tensor(w, t), dtype=np.uint8).copy()
    want = np.zeros((n, rows), dtype=np.float32)
    checkers.oracle_lib().bo_mul_mat(t, wb, x, want, k, rows, n)
    got = capi.op_mul_mat_tc(t, wb, x, rows)
    err = np.abs(got - want).max()
    tol = 3e-5 * np.abs(want).max() + 1e-6
    assert err <= tol, f"{name} {shape}: max|d|={err:.3e} tol={tol:.3e}"
    # and it must agree with the exact-order GPU kernel to the same tolerance
    exact = capi.op_mul_mat(t, wb, x, rows)
    assert np.abs(got - exact).max() <= tol


@pytest.mark.parametrize("name", ["q4_0", "q4_1", "q5_0", "q5_1", "q8_0"])
@pytest.mark.parametrize("shape", [(1024, 256, 64), (4096, 136, 70), (1024, 1000, 200), (128, 128, 33), (64, 8, 5), (1024, 3072, 17)])
def test_mul_mat_tensor_core_exact(checkers, capi, name, shape):
    """the bit-exact tcgen05 matmul (csrc/bgpt_tc.cuh, k_gemm_tc_x): every token becomes 8 masked activation columns, the MMA
    delivers the reference's 8 four-element partial sums per block as exact int32, the epilogue runs the 8 fma chains in block order
    -- the result must be the oracle's bits for every format, ragged rows / tokens and K = 64 .. 4096"""
    k, rows, n = shape
    t = TYPES[name]
    rng = np.random.default_rng(k + rows * 3 + n + t)
    w = (rng.standard_normal((rows, k)) * 0.02).astype(np.float32)
    x = rng.standard_normal((n, k)).astype(np.float32)
    x[0, :32] = 0.0                                       # an all-zero activation block
    wb = np.frombuffer(gf.encode_tensor(w, t), dtype=np.uint8).copy()
    want = np.zeros((n, rows), dtype=np.float32)
    checkers.oracle_lib().bo_mul_mat(t, wb, x, want, k, rows, n)
    got = capi.op_mul_mat_tcx(t, wb, x, rows)
    bad = np.flatnonzero(got.view(np.uint32).ravel() != want.view(np.uint32).ravel())
    assert bad.size == 0, f"{name} {shape}: {bad.size}/{got.size} outputs differ, max|d|={np.abs(got - want).max():.3e}"


@pytest.mark.parametrize("name", ["q4_0", "q4_1", "q5_0", "q5_1", "q8_0"])
@pytest.mark.parametrize("shape", [(1024, 256, 64), (4096, 128, 70), (1024, 1024, 200), (64, 128, 5), (128, 128, 33), (1024, 3072, 17),
                                   (1024, 2560, 160)])
def test_mul_mat_tensor_core_exact_tma_fed(checkers, capi, name, shape):
    """the warp-specialised, TMA-fed form of the bit-exact tcgen05 matmul (csrc/bgpt_tcw.cuh, k_tcw_exact: operands from the decoded
    weight planes and the expanded activations by cp.async.bulk.tensor, double-buffered TMEM, FFMA2 chains) -- the oracle's bits for
    every format, ragged token counts, one to many tiles per CTA and K = 64 .. 4096"""
    k, rows, n = shape
    t = TYPES[name]
    rng = np.random.default_rng(k + rows * 3 + n + t)
    w = (rng.standard_normal((rows, k)) * 0.02).astype(np.float32)
    x = rng.standard_normal((n, k)).astype(np.float32)
    x[0, :32] = 0.0                                       # an all-zero activation block
    wb = np.frombuffer(gf.encode_tensor(w, t), dtype=np.uint8).copy()
    want = np.zeros((n, rows), dtype=np.float32)
    checkers.oracle_lib().bo_mul_mat(t, wb, x, want, k, rows, n)
    got = capi.op_mul_mat_tcw(t, wb, x, rows)
    bad = np.flatnonzero(got.view(np.uint32).ravel() != want.view(np.uint32).ravel())
    assert bad.size == 0, f"{name} {shape}: {bad.size}/{got.size} outputs differ, max|d|={np.abs(got - want).max():.3e}, first={bad[:6]}"


@pytest.mark.parametrize("shape", [(1024, 256, 64), (4096, 128, 130), (256, 128, 5), (1024, 3072, 300), (1024, 1024, 128)])
def test_mul_mat_tensor_core_f16(checkers, capi, shape):
    """F16 weights on the K-accumulating tcgen05 matmul (csrc/bgpt_tcw.cuh, k_tcw_f16: weights by TMA where they lie, fp16-rounded
    activations, f32 accumulation in TMEM).  Same products as ggml_vec_dot_f16 (ggml.c:2409-2443), added in the tensor core's order:
    tolerance 2e-5 of the largest output (the f32 summation-order noise of a K <= 4096 dot), not bit equality."""
    k, rows, n = shape
    rng = np.random.default_rng(k + rows * 5 + n)
    w = (rng.standard_normal((rows, k)) * 0.02).astype(np.float32)
    x = rng.standard_normal((n, k)).astype(np.float32)
    wb = np.frombuffer(gf.encode_tensor(w, 1), dtype=np.uint8).copy()
    want = np.zeros((n, rows), dtype=np.float32)
    checkers.oracle_lib().bo_mul_mat(1, wb, x, want, k, rows, n)
    got = capi.op_mul_mat_tcw(1, wb, x, rows)
    err = np.abs(got - want).max()
    tol = 2e-5 * np.abs(want).max() + 1e-7
    print(f"f16 tcgen05 matmul {shape}: max|d| = {err:.3e} = {err / np.abs(want).max():.2e} of the largest output; mean signed d = {np.mean(got - want):.2e}")
    assert err <= tol, f"f16 {shape}: max|d|={err:.3e} tol={tol:.3e}"


@pytest.mark.parametrize("name", ["q4_0", "q4_1", "q5_0", "q5_1", "q8_0"])
def test_quantize_weights_equals_reference_quantiser(capi, name):
    """device f32 -> Qx blocks (csrc/bgpt_quant.cuh) == quantize_row_q*_reference as the reference's `quantize` tool
    runs it (ggml.c:892-1094).  The checker is ggml_file's numpy restatement, itself pinned to the tool's output and to
    ggml's own test signal (tests/test_oracle_golden.py).  Cases: ggml's 0.1 + 2 cos(i) signal, synthetic weights, an
    all-zero block, ties of |max| with opposite signs (first occurrence wins), one block, a ragged warp (33 blocks)."""
    t = TYPES[name]
    rng = np.random.default_rng(99 + t)
    cases = []
    n = 32 * 128
    cases.append((0.1 + 2.0 * np.cos(np.arange(n, dtype=np.float32))).astype(np.float32))          # ggml/tests/test-quantize-fns.cpp:26-30
    w = (rng.standard_normal(32 * 4097) * 0.02).astype(np.float32)
    w[64:96] = 0.0
    w[96] = 0.5; w[97] = -0.5                                     # equal magnitudes: the first one sets the sign of d
    w[128] = -0.25; w[140] = 0.25
    cases.append(w)
    cases.append((rng.standard_normal(32) * 3).astype(np.float32))
    cases.append((rng.standard_normal(32 * 33) * 1e-3).astype(np.float32))
    cases.append(np.linspace(-4, 4, 32 * 7, dtype=np.float32))
    for i, x in enumerate(cases):
        got = capi.op_quantize_weights(t, x)
        want = gf.QUANTIZERS[t](x)
        bad = np.flatnonzero(got != want)
        assert bad.size == 0, f"{name} case {i}: {bad.size}/{got.size} bytes differ, first at {bad[:8]} (block {bad[0] // gf.TYPE_SIZE[t]})"


def test_device_topk_selects_what_partial_sort_selects(capi, zoo):
    """bgpt_cuda_eval_topk: K (logit, id) pairs sorted by logit descending == the K best of the full logit row (stable order is
    irrelevant: ties are flagged `exact = False` and come with the full row)"""
    M = capi.Model.load(zoo.path("small", "q8_0"))
    toks = gf.synth_tokens(12, gf.SMALL.n_vocab, seed=3)
    for t, n_past in ((toks[:8], 0), (toks[8:9], 8), (toks[9:10], 9)):
        full = M.eval(t, n_past)
        for k in (1, 5, 40, 128):
            vals, ids, exact, fb = M.eval_topk(t, n_past, k)
            order = np.argsort(-full, kind="stable")[:k]
            if exact:
                assert ids.tolist() == order.tolist() and np.array_equal(vals.view(np.uint32), full[order].view(np.uint32)), (n_past, k)
                assert fb is None
            else:                                          # equal logits somewhere in the top K: the full row must have come along
                assert fb is not None and np.array_equal(fb.view(np.uint32), full.view(np.uint32))
            assert np.array_equal(np.sort(vals)[::-1], vals)
    M.close()


def test_device_topk_after_the_persistent_kernel_full_vocabulary(capi, zoo):
    """the full-size vocabulary (42384): a single-token step on the generation-5 kernel selects in the kernel's tail (the per-CTA maxima
    give a threshold, the last CTA collects the logits above it and ranks them), on generation 4 in two launches behind the kernel
    (k_topk_filter + one ranking CTA), a prompt batch in three (slices of 256 -> groups of 16 slices -> one CTA); all must return the
    K best of the full row"""
    M = capi.Model.load(zoo.path("base", "q4_0"))
    toks = gf.synth_tokens(12, gf.BASE.n_vocab, seed=31)
    # decode path 1: generation 5 -- the selection is the kernel's tail (topk_tail); 3: generation 4 -- filter + ranking launches
    for path in (1, 3):
        M.set_decode_path(path)
        for t, n_past in ((toks[:8], 0), (toks[8:9], 8), (toks[9:10], 9), (toks[10:11], 10)):
            full = M.eval(t, n_past).copy()
            for k in (1, 5, 40, 128):
                vals, ids, exact, fb = M.eval_topk(t, n_past, k)
                order = np.argsort(-full, kind="stable")[:k]
                if exact:
                    assert ids.tolist() == order.tolist() and np.array_equal(vals.view(np.uint32), full[order].view(np.uint32)), (path, n_past, k)
                else:
                    assert fb is not None and np.array_equal(fb.view(np.uint32), full.view(np.uint32))
    M.close()


@pytest.mark.parametrize("ftype", ["q4_0", "q5_1", "f16"])
def test_chained_launches_return_what_single_launches_return(capi, zoo, ftype):
    """bgpt_cuda_eval_topk queues the kernel of position p + 1 while the call for p waits (bgpt_cuda_set_chain): the sampling loop must
    see the same (logit, id) pairs with and without it, through withdrawals (another call in between, a jump in n_past, another k),
    through a caller that is slower than the kernel's patience (the queued kernel gives up, a fresh launch serves the call) and up to
    the last position of the context"""
    import time
    M = capi.Model.load(zoo.path("base", ftype))
    if M.decode_generation != 5:
        M.close(); pytest.skip("chained launches need the generation-5 decode kernel")
    n_pos = gf.BASE.n_positions
    M.decode_greedy(2, 0, n_pos)                                        # every position of the KV cache holds defined values

    def run(chain, plan):
        M.set_chain(chain)
        out = []
        tok = np.array([2], dtype=np.int32)
        for what, p, k in plan:
            if what == "full":
                full = M.eval(tok, p)
                out.append(("full", p, int(np.argmax(full))))
                continue
            if what == "sleep":
                time.sleep(0.02)
                continue
            vals, ids, exact, fb = M.eval_topk(tok, p, k)
            out.append((p, k, vals.copy().view(np.uint32).tolist(), ids.copy().tolist(), exact))
            tok[0] = int(ids[0]) if exact else int(np.argmax(fb))
        return out

    plan = [("topk", p, 40) for p in range(0, 24)]                     # a plain sampling loop
    plan += [("full", 24, 0)] + [("topk", p, 40) for p in range(24, 30)]   # another call withdraws the queued kernel
    plan += [("topk", p, 5) for p in range(30, 34)]                    # another k
    plan += [("topk", 12, 40), ("topk", 13, 40)]                       # back in the context
    plan += [("sleep", 0, 0), ("topk", 14, 40), ("sleep", 0, 0), ("topk", 15, 40), ("topk", 16, 40)]   # slower than the queued kernel's patience
    plan += [("topk", p, 40) for p in range(n_pos - 3, n_pos)]         # the last positions: nothing is queued behind the last one
    want = run(0, plan)
    got = run(1, plan)
    assert got == want
    got2 = run(-1, plan)                                               # the default (on unless BGPT_CHAIN=0)
    assert got2 == want
    M.close()


@pytest.mark.parametrize("n", [1000, 4096, 42384, 50001])
def test_device_topk_selection_and_tie_flags(capi, n):
    """the selection kernel(s) on crafted logit rows: n >= 4096 runs the two-launch form (every 256-logit slice keeps its K + 1
    best, one CTA selects among the candidates), below that the single-CTA radix select.  Distinct values: ids == stable argsort.
    Ties that make std::partial_sort's choice or order ambiguous (biogpt.cpp:908-980) must be flagged: the K-th value equal to the
    (K+1)-th in another slice, in the same slice, more than K + 1 equal values inside one slice, two equal values inside the top K;
    ties strictly below the K-th value must NOT be flagged."""
    rng = np.random.default_rng(n)
    base = rng.permutation(n).astype(np.float32) * 0.25 - 1000.0          # all distinct
    for k in (1, 5, 40, 128):
        order = np.argsort(-base, kind="stable")
        vals, ids, exact = capi.op_topk(base, k)
        assert exact and ids.tolist() == order[:k].tolist() and np.array_equal(vals, base[order[:k]]), (n, k)
        kth, nxt = order[k - 1], order[k]
        # boundary tie: the (K+1)-th gets the K-th value -- once far away (another slice), once right next to it
        for where in ((int(kth) + n // 2) % n, (int(kth) + 1) % n):
            x = base.copy()
            if where in order[:k].tolist():
                continue
            x[where] = x[kth]
            _, _, exact = capi.op_topk(x, k)
            assert not exact, (n, k, where)
        # more than K + 1 equal values at the threshold inside one slice
        x = base.copy()
        s0 = (int(kth) // 256) * 256
        idx = [i for i in range(s0, min(s0 + 256, n)) if i not in set(order[:k - 1].tolist())][:k + 3]
        x[idx] = x[kth]
        _, _, exact = capi.op_topk(x, k)
        assert not exact, (n, k, "slice full of ties")
        if k >= 2:                                                   # duplicate inside the top K
            x = base.copy(); x[order[0]] = x[order[1]]
            _, _, exact = capi.op_topk(x, k)
            assert not exact, (n, k, "duplicate in the top K")
        # ties below the K-th value are harmless
        x = base.copy(); x[order[k + 5]] = x[order[k + 6]] if k + 6 < n else x[order[k + 5]]
        x[nxt] = np.float32(x[kth] - 0.125)
        vals, ids, exact = capi.op_topk(x, k)
        want = np.argsort(-x, kind="stable")[:k]
        assert exact and ids.tolist() == want.tolist(), (n, k, "ties below the threshold")

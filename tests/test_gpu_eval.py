"""GPU tier: whole `biogpt_eval` steps through the C ABI against the oracle and the committed
reference logits -- bit for bit, every format, prompt batches (un-masked) and decode steps."""
import os

import numpy as np
import pytest

from conftest import FTYPES, ROOT, gf

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _diff(tag, got, want):
    bad = np.flatnonzero(_bits(got).ravel() != _bits(want).ravel())
    return f"{tag}: {bad.size}/{got.size} logits differ, max|d|={np.abs(got - want).max():.3e}"


@pytest.mark.parametrize("ftype", FTYPES)
def test_tiny_matches_reference_golden(capi, zoo, ftype):
    """the committed logits came from the UNMODIFIED reference (tests/golden/make_golden.py)"""
    gold = np.load(os.path.join(ROOT, "tests", "golden", "tiny_logits.npz"))
    M = capi.Model.load(zoo.path("tiny", ftype))
    toks, sched, want = gold["tokens"], gold["schedule"], gold[f"logits_{ftype}"]
    pos = 0
    for i, (n, n_past) in enumerate(sched):
        got = M.eval(toks[pos:pos + n], int(n_past))
        pos += n
        assert np.array_equal(_bits(got), _bits(want[i])), _diff(f"{ftype} eval {i} (n={n}, n_past={n_past})", got, want[i])
    M.close()


@pytest.mark.parametrize("ftype", FTYPES)
def test_small_matches_oracle_with_taps(checkers, capi, zoo, ftype):
    """d_model 256, d_kv 64, d_ff 1024, 3 layers; intermediate taps localise any mismatch"""
    p = zoo.path("small", ftype)
    O = checkers.Oracle(p)
    M = capi.Model.load(p, max_batch=16)
    toks = gf.synth_tokens(70, gf.SMALL.n_vocab, seed=21)
    pos = 0
    for n in (8, 8, 1, 1, 5, 1, 16, 1, 1, 9, 1):
        want, wt = O.eval(toks[pos:pos + n], pos, taps=True)
        got, gt = M.eval(toks[pos:pos + n], pos, taps=True)
        for a, b in (("embed", "embed"), ("layer0_q", "layer0_q"), ("layer0_att", "layer0_att"), ("layer0_out", "layer0_out")):
            assert np.array_equal(_bits(gt[a]), _bits(wt[b])), _diff(f"{ftype} tap {a} at pos {pos} n={n}", gt[a], wt[b])
        assert np.array_equal(_bits(got), _bits(want)), _diff(f"{ftype} logits at pos {pos} n={n}", got, want)
        pos += n
    O.close(); M.close()


@pytest.mark.parametrize("ftype", ["f16", "q4_0", "q8_0"])
def test_greedy_ids_identical(checkers, capi, zoo, ftype):
    """device-side greedy loop (argmax on the GPU, no host round trip) == oracle greedy ids"""
    p = zoo.path("small", ftype)
    O = checkers.Oracle(p)
    M = capi.Model.load(p)
    prompt = gf.synth_tokens(6, gf.SMALL.n_vocab, seed=2)
    lo = O.eval(prompt, 0)
    lg = M.eval(prompt, 0)
    assert np.array_equal(_bits(lo), _bits(lg))
    first = int(np.argmax(lo))
    steps = 48
    ids, ms = M.decode_greedy(first, len(prompt), steps)
    want = []
    tok = first
    for i in range(steps):
        l = O.eval(np.array([tok], np.int32), len(prompt) + i)
        tok = int(np.argmax(l))
        want.append(tok)
    assert ids.tolist() == want
    O.close(); M.close()


@pytest.mark.parametrize("ftype", ["q5_1", "f16"])
def test_lockstep_streams_equal_single_stream(capi, zoo, ftype):
    """config 4: S sequences decoded in lock step share each weight read; every stream must equal
    its own single-stream run bit for bit"""
    p = zoo.path("small", ftype)
    S, steps = 5, 12
    seqs = [gf.synth_tokens(steps, gf.SMALL.n_vocab, seed=100 + s) for s in range(S)]
    single = []
    M = capi.Model.load(p)
    for s in range(S):
        single.append(np.stack([M.eval(seqs[s][i:i + 1], i) for i in range(steps)]))
    M.set_streams(S)
    for i in range(steps):
        out = M.eval_streams(np.array([seqs[s][i] for s in range(S)], np.int32), i)
        for s in range(S):
            assert np.array_equal(_bits(out[s]), _bits(single[s][i])), (ftype, s, i)
    M.close()


@pytest.mark.parametrize("size,ftype,S", [("small", "q8_0", 5), ("base", "q5_1", 8), ("tiny", "f16", 3)])
def test_lockstep_streams_topk_equals_full_rows(capi, zoo, size, ftype, S):
    """bgpt_cuda_eval_streams_topk: per stream the k largest (logit, id) pairs by logit descending == the k best of the row
    bgpt_cuda_eval_streams returns (ties are flagged exact = 0 and come with the full row)"""
    hp = {"small": gf.SMALL, "base": gf.BASE, "tiny": gf.TINY}[size]
    M = capi.Model.load(zoo.path(size, ftype))
    M.set_streams(S)
    steps = 4
    seqs = [gf.synth_tokens(steps, hp.n_vocab, seed=300 + s) for s in range(S)]
    for i in range(steps):
        toks = np.array([seqs[s][i] for s in range(S)], np.int32)
        full = M.eval_streams(toks, i).copy()
        for k in (1, 40, 128):
            if k > hp.n_vocab:
                continue
            vals, ids, n_out, exact, fb = M.eval_streams_topk(toks, i, k)
            for s in range(S):
                assert n_out[s] == k
                order = np.argsort(-full[s], kind="stable")[:k]
                if exact[s]:
                    assert ids[s].tolist() == order.tolist() and np.array_equal(_bits(vals[s]), _bits(full[s][order])), (size, ftype, s, i, k)
                else:
                    assert np.array_equal(_bits(fb[s]), _bits(full[s])), (size, ftype, s, i, k)
    M.close()


@pytest.mark.parametrize("ftype", ["q4_0", "f16", "q8_0", "q5_1"])
def test_base_shape_matches_oracle(checkers, capi, zoo, ftype):
    """true BioGPT-base shapes (d_model 1024, 24 layers, 16 heads, d_ff 4096, vocab 42384),
    synthetic weights: an 8-token un-masked prompt batch, then decode steps"""
    p = zoo.path("base", ftype)
    O = checkers.Oracle(p)
    M = capi.Model.load(p)
    toks = gf.synth_tokens(12, gf.BASE.n_vocab, seed=9)
    want = O.eval(toks[:8], 0); got = M.eval(toks[:8], 0)
    assert np.array_equal(_bits(got), _bits(want)), _diff(f"{ftype} prompt", got, want)
    for i in range(8, 12):
        want = O.eval(toks[i:i + 1], i); got = M.eval(toks[i:i + 1], i)
        assert np.array_equal(_bits(got), _bits(want)), _diff(f"{ftype} decode {i}", got, want)
    O.close(); M.close()


@pytest.mark.parametrize("ftype", ["f16", "q4_0", "q4_1", "q5_0", "q5_1", "q8_0"])
@pytest.mark.parametrize("size", ["tiny", "small"])
def test_persistent_kernel_equals_oracle_and_per_op_path(checkers, capi, zoo, size, ftype):
    """single-token steps run on ONE persistent kernel (grid barriers between phases); it must
    give the oracle's bits and the per-operator schedule's bits at every position, including
    T crossing the 32-wide boundary and the scalar tails (T % 32 in 1..3 and 4..31)"""
    hp = {"tiny": gf.TINY, "small": gf.SMALL}[size]
    p = zoo.path(size, ftype)
    O = checkers.Oracle(p)
    M = capi.Model.load(p)
    assert M.decode_path == 1, "persistent kernel not available: " + capi.last_error()
    n_steps = min(hp.n_positions, 70)
    toks = gf.synth_tokens(n_steps, hp.n_vocab, seed=33)
    got1 = []
    for i in range(n_steps):
        want = O.eval(toks[i:i + 1], i)
        got = M.eval(toks[i:i + 1], i)
        assert np.array_equal(_bits(got), _bits(want)), _diff(f"{size}/{ftype} persistent kernel at n_past={i}", got, want)
        got1.append(got)
    M.set_decode_path(0)
    for i in range(n_steps):
        got = M.eval(toks[i:i + 1], i)
        assert np.array_equal(_bits(got), _bits(got1[i])), (size, ftype, i)
    O.close(); M.close()


@pytest.mark.parametrize("ftype", ["q4_0", "q4_1", "q5_0", "q5_1", "q8_0", "f16"])
def test_persistent_generations_equal_oracle_and_each_other(checkers, capi, zoo, ftype):
    """BioGPT-base layer shapes: single-token steps run on the generation-5 persistent kernel (clusters of 4 CTAs, one
    attention head per cluster, DSMEM exchange inside the head, chain-order dot products, TMA weight ring).  Bits must equal
    the oracle's, generation 4's (tagged-word exchange, no clusters) and generation 3's (grid barriers) at every position
    class: T = 1, the 32-wide boundary and both scalar tails, the second K pass (T > 512) and the end of the context; the
    vocabulary (3001 rows) leaves ragged lm_head tiles.  F16 weights run the same kernel with fp16 records and one row per
    warp (lane = running sum, ggml_vec_dot_f16's 32 lanes)."""
    hp = gf.NARROW
    p = zoo.path("narrow", ftype)
    O = checkers.Oracle(p)
    M = capi.Model.load(p, max_batch=64)
    assert M.decode_generation == 5, "generation-5 kernel not selected: " + capi.last_error()
    toks = gf.synth_tokens(hp.n_positions, hp.n_vocab, seed=77)
    windows = [(0, 70), (500, 531), (990, 1024)]
    pos = 0
    got5 = {}
    for lo, hi in windows:
        while pos < lo:                                   # fill the cache with un-masked prompt batches (skinny-batch schedule)
            n = min(16, lo - pos)
            want = O.eval(toks[pos:pos + n], pos); got = M.eval(toks[pos:pos + n], pos)
            assert np.array_equal(_bits(got), _bits(want)), _diff(f"{ftype} prompt at {pos}", got, want)
            pos += n
        for i in range(lo, hi):
            want = O.eval(toks[i:i + 1], i)
            got = M.eval(toks[i:i + 1], i)
            assert np.array_equal(_bits(got), _bits(want)), _diff(f"{ftype} generation 5 at n_past={i}", got, want)
            got5[i] = got
        pos = hi
    for path, gen in ((3, 4 if ftype != "f16" else 3), (2, 3)):       # F16: generations 5 and 3 (generation 4 is quantised-only)
        M.set_decode_path(path)
        assert M.decode_generation == gen
        for lo, hi in windows:
            for i in range(lo, hi):                       # the cache rows are already there: same evals again
                got = M.eval(toks[i:i + 1], i)
                assert np.array_equal(_bits(got), _bits(got5[i])), (ftype, gen, i)
    ids = {}
    for path in (1, 3, 2):
        M.set_decode_path(path)
        ids[path], _ = M.decode_greedy(int(toks[0]), 0, 40)
    assert ids[1].tolist() == ids[3].tolist() == ids[2].tolist()
    O.close(); M.close()


@pytest.mark.parametrize("ftype", ["q4_0", "q4_1", "q5_0", "q5_1", "q8_0"])
def test_skinny_schedule_equals_oracle_and_per_op_path(checkers, capi, zoo, ftype):
    """BioGPT-base layer shapes, 2..111 token rows: the fused skinny-batch schedule (csrc/bgpt_skinny.cuh: LayerNorm /
    quantise / GELU folded into warp-per-row matmul kernels, attention with a quantising epilogue, programmatic dependent
    launch).  Un-masked prompt batches of every tile shape (4- and 8-row tiles, two tiles, ragged last tile) at T across
    the 32-wide boundary, both scalar tails, T > 512 and the end of the context must give the oracle's bits and the
    per-operator schedule's bits."""
    hp = gf.NARROW
    p = zoo.path("narrow", ftype)
    O = checkers.Oracle(p)
    M = capi.Model.load(p, max_batch=128)
    M.set_batch_path(2)
    assert M.batch_path(8) == 2 and M.batch_path(2) == 2 and M.batch_path(9) == 1, capi.last_error()    # opt-in, 2..8 rows: bgpt_rows.cuh
    M.set_batch_path(1)
    assert M.batch_path(8) == 1 and M.batch_path(1) == 0 and M.batch_path(111) == 1 and M.batch_path(128) == 0, capi.last_error()
    toks = gf.synth_tokens(hp.n_positions, hp.n_vocab, seed=55)
    sizes = [8, 8, 5, 2, 3, 4, 7, 16, 9, 1, 8, 31, 8, 6]             # 116 positions: T = 8, 16, 21, 23, 26, 30, 37, 53, 62, 63, 71, 102, ...
    sched, pos = [], 0
    for n in sizes:
        sched.append((pos, n)); pos += n
    while pos < 500:
        sched.append((pos, 24)); pos += 24                              # 3 tiles of 8
    for n in (8, 8, 8, 3, 8, 8, 40, 64, 100, 111):                      # many tiles per launch: the range the tensor-core path used to serve
        sched.append((pos, n)); pos += n
    while pos < 1000:
        sched.append((pos, 16)); pos += 16
    while pos < hp.n_positions:
        n = min(8, hp.n_positions - pos); sched.append((pos, n)); pos += n
    got_fused = []
    for pos, n in sched:
        want = O.eval(toks[pos:pos + n], pos)
        got = M.eval(toks[pos:pos + n], pos)
        assert np.array_equal(_bits(got), _bits(want)), _diff(f"{ftype} fused schedule n={n} at {pos}", got, want)
        got_fused.append(got)
    M.set_batch_path(0)
    assert M.batch_path(8) == 0
    for (pos, n), ref in zip(sched, got_fused):
        got = M.eval(toks[pos:pos + n], pos)
        assert np.array_equal(_bits(got), _bits(ref)), (ftype, pos, n)
    O.close(); M.close()


@pytest.mark.parametrize("ftype", ["q5_1", "q4_0", "q8_0"])
def test_skinny_lockstep_streams_equal_single_stream(capi, zoo, ftype):
    """config 4 at BioGPT-base layer shapes: S sequences in lock step on the fused skinny-batch schedule; every stream must
    equal its own single-stream run (generation-4 persistent kernel) bit for bit, for 4- and 8-row tiles and two tiles"""
    hp = gf.NARROW
    p = zoo.path("narrow", ftype)
    steps = 40
    for S in (8, 3, 12):
        seqs = [gf.synth_tokens(steps, hp.n_vocab, seed=300 + 7 * S + s) for s in range(S)]
        M = capi.Model.load(p, max_batch=16)
        single = [np.stack([M.eval(seqs[s][i:i + 1], i) for i in range(steps)]) for s in range(S)]
        M.set_streams(S)
        M.set_batch_path(1)
        assert M.batch_path(S) == 1
        for i in range(steps):
            out = M.eval_streams(np.array([seqs[s][i] for s in range(S)], np.int32), i)
            for s in range(S):
                assert np.array_equal(_bits(out[s]), _bits(single[s][i])), _diff(f"{ftype} S={S} stream {s} step {i}", out[s], single[s][i])
        M.close()


@pytest.mark.parametrize("ftype", ["q4_0", "q4_1", "q5_0", "q5_1", "q8_0"])
def test_multi_row_kernel_equals_oracle_and_skinny_schedule(checkers, capi, zoo, ftype):
    """BioGPT-base layer shapes, 2..8 token rows on the persistent multi-row kernel (csrc/bgpt_rows.cuh: one launch per eval,
    stage boundaries are counters in L2).  Un-masked prompt batches of every size 2..8 at T across the 32-wide boundary, both
    scalar tails of the V product, T > 512 and the end of the context must give the oracle's bits; the KV cache the kernel
    wrote must serve the single-token persistent kernel and the skinny schedule (and theirs must serve it)."""
    hp = gf.NARROW
    p = zoo.path("narrow", ftype)
    O = checkers.Oracle(p)
    M = capi.Model.load(p, max_batch=16)
    assert [M.eval_path(n) for n in (1, 2, 8, 9)] == [3, 1, 1, 1], capi.last_error()      # the default: fused skinny-batch schedule
    M.set_batch_path(2)
    assert [M.eval_path(n) for n in (1, 2, 8, 9)] == [3, 5, 5, 1], capi.last_error()
    toks = gf.synth_tokens(hp.n_positions, hp.n_vocab, seed=4242)
    sizes = [8, 8, 5, 2, 3, 4, 7, 6, 1, 8, 2, 8, 12, 3, 8]            # 1 row: decode kernel, 12 rows: skinny schedule, on the same cache
    sched, pos = [], 0
    for n in sizes:
        sched.append((pos, n)); pos += n
    while pos < 520:
        sched.append((pos, 8)); pos += 8
    for n in (7, 5, 3, 2, 4, 6):
        sched.append((pos, n)); pos += n
    while pos < hp.n_positions:
        n = min(8, hp.n_positions - pos); sched.append((pos, n)); pos += n
    got_rows = []
    for pos, n in sched:
        want = O.eval(toks[pos:pos + n], pos)
        got = M.eval(toks[pos:pos + n], pos)
        assert np.array_equal(_bits(got), _bits(want)), _diff(f"{ftype} multi-row kernel n={n} at {pos} (path {M.eval_path(n)})", got, want)
        got_rows.append(got)
    M.set_batch_path(1)
    assert M.eval_path(8) == 1
    for (pos, n), ref in zip(sched, got_rows):
        got = M.eval(toks[pos:pos + n], pos)
        assert np.array_equal(_bits(got), _bits(ref)), (ftype, pos, n)
    O.close(); M.close()


@pytest.mark.parametrize("ftype", ["q5_1", "q4_0", "q8_0"])
def test_multi_row_kernel_lockstep_streams_equal_single_stream(capi, zoo, ftype):
    """BASELINE configs[3] per GPU at BioGPT-base layer shapes: S <= 8 sequences in lock step on the persistent multi-row kernel;
    every stream must equal its own single-stream run (persistent decode kernel) bit for bit -- through the host-buffer call and
    through the device-side greedy loop"""
    hp = gf.NARROW
    p = zoo.path("narrow", ftype)
    steps = 40
    for S in (8, 2, 5):
        seqs = [gf.synth_tokens(steps, hp.n_vocab, seed=900 + 7 * S + s) for s in range(S)]
        M = capi.Model.load(p, max_batch=16)
        single = [np.stack([M.eval(seqs[s][i:i + 1], i) for i in range(steps)]) for s in range(S)]
        M.set_streams(S)
        M.set_batch_path(2)
        assert M.batch_path(S) == 2 and M.eval_path(S) == 5
        for i in range(steps):
            out = M.eval_streams(np.array([seqs[s][i] for s in range(S)], np.int32), i)
            for s in range(S):
                assert np.array_equal(_bits(out[s]), _bits(single[s][i])), _diff(f"{ftype} S={S} stream {s} step {i}", out[s], single[s][i])
        first = [int(seqs[s][0]) for s in range(S)]
        ids_rows, _ = M.decode_greedy_streams(first, 0, 24)
        M.set_batch_path(1)
        ids_sk, _ = M.decode_greedy_streams(first, 0, 24)
        assert np.array_equal(np.asarray(ids_rows), np.asarray(ids_sk)), (ftype, S)
        M.close()


def _fast_checker(checkers, path, n_batch=8):
    """the unmodified reference on all host threads (AVX2) where oracle/_ref was built, else the plain-C restatement;
    tests/test_oracle_vs_ref.py pins the two to each other bit for bit"""
    return checkers.Ref(path, n_batch=n_batch) if checkers.have_ref() else checkers.Oracle(path)


def _top5(v):
    return set(np.argsort(-v, kind="stable")[:5].tolist())


@pytest.mark.parametrize("ftype", ["q4_0", "q4_1", "q5_0", "q5_1", "q8_0"])
def test_large_prompt_batches_are_bit_exact(checkers, capi, zoo, ftype):
    """whole evals of 112 / 128 / 256 / 1024 rows (BioGPT-base layer shapes, 2 layers) on their DEFAULT paths -- the skinny-batch
    schedule below 128 rows, the per-operator schedule with the bit-exact tcgen05 matmul (k_gemm_tc_x: the reference's 8 running
    sums per row through masked activation columns) from 128 rows on: logits bit-identical to the reference, a 24-token greedy
    continuation on the KV cache that pass wrote gives bit-identical logits, and the skinny-batch schedule gives the same bits
    at every size"""
    hp = gf.NARROW
    p = zoo.path("narrow", ftype)
    toks = gf.synth_tokens(hp.n_positions, hp.n_vocab, seed=808)
    for rows in (112, 128, 256, 1024):
        R = _fast_checker(checkers, p, n_batch=rows)
        M = capi.Model.load(p, max_batch=rows)
        assert M.eval_path(rows) == (4 if rows >= 128 else 1) and M.eval_path(1) == 3, capi.last_error()
        want = R.eval(toks[:rows], 0)
        got = M.eval(toks[:rows], 0)
        assert np.array_equal(_bits(got), _bits(want)), _diff(f"{ftype} rows={rows} (path {M.eval_path(rows)})", got, want)
        if rows >= 128:
            M.set_tcw(0)                                                   # the older single-stage form of the same matmul (k_gemm_tc_xf)
            got = M.eval(toks[:rows], 0)
            assert np.array_equal(_bits(got), _bits(want)), _diff(f"{ftype} rows={rows} (k_gemm_tc_xf)", got, want)
            M.set_tcw(1)
            M.set_tcx_min_rows(0)
            assert M.eval_path(rows) == 1
            got = M.eval(toks[:rows], 0)
            assert np.array_equal(_bits(got), _bits(want)), _diff(f"{ftype} rows={rows} (skinny-batch schedule)", got, want)
            M.set_tcx_min_rows(128)
        if rows < hp.n_positions:
            tok = int(np.argmax(want))
            for i in range(24):
                lr = R.eval(np.array([tok], np.int32), rows + i)
                lm = M.eval(np.array([tok], np.int32), rows + i)
                assert np.array_equal(_bits(lm), _bits(lr)), _diff(f"{ftype} rows={rows} continuation step {i}", lm, lr)
                tok = int(np.argmax(lr))
        R.close(); M.close()


@pytest.mark.parametrize("ftype,rows", [("q4_0", 128), ("q8_0", 1024)])
def test_large_prompt_batch_base_model_bit_exact(checkers, capi, zoo, ftype, rows):
    """the same on the full 24-layer, 42384-row-vocabulary model"""
    hp = gf.BASE
    p = zoo.path("base", ftype)
    toks = gf.synth_tokens(hp.n_positions, hp.n_vocab, seed=909)
    R = _fast_checker(checkers, p, n_batch=rows)
    M = capi.Model.load(p, max_batch=rows)
    assert M.eval_path(rows) == 4
    want = R.eval(toks[:rows], 0)
    got = M.eval(toks[:rows], 0)
    assert np.array_equal(_bits(got), _bits(want)), _diff(f"base {ftype} rows={rows}", got, want)
    R.close(); M.close()


# The integer tcgen05 matmul (csrc/bgpt_tc.cuh) adds ONE f32 term per 32-block instead of the reference's 8 running sums.
# Each matmul is within 3e-5 relative (tests/test_gpu_ops.py) but activations are re-quantised to int8 in front of every
# matmul, so the drift compounds: measured max|dlogit| 1.7e-2 (2 layers) .. 5.3e-2 (24 layers) on logits of spread ~4, and
# top-5 sets of a flat synthetic distribution are not stable.  That fails the north_star gate for Qx, so the path is OPT-IN
# (bgpt_cuda_set_tc_min_rows) and this test only pins its documented envelope: same argmax, max|dlogit| <= 1e-1.
TC_LOGIT_TOL = 1e-1


@pytest.mark.parametrize("ftype", ["q4_0", "q5_1", "q8_0"])
def test_opt_in_integer_tensor_core_path_envelope(checkers, capi, zoo, ftype):
    hp = gf.NARROW
    p = zoo.path("narrow", ftype)
    toks = gf.synth_tokens(hp.n_positions, hp.n_vocab, seed=808)
    for rows in (128, 1024):
        R = _fast_checker(checkers, p, n_batch=rows)
        M = capi.Model.load(p, max_batch=rows)
        M.set_tc_min_rows(112)
        assert M.eval_path(rows) == 2 and M.eval_path(111) == 1
        want = R.eval(toks[:rows], 0)
        got = M.eval(toks[:rows], 0)
        err = float(np.abs(got - want).max())
        print(f"opt-in integer tcgen05 path {ftype} rows={rows}: max|dlogit| = {err:.3e}")
        assert int(np.argmax(got)) == int(np.argmax(want)) and err <= TC_LOGIT_TOL, f"{ftype} rows={rows}: max|dlogit|={err:.3e}"
        M.set_tc_min_rows(0)
        assert M.eval_path(rows) == 4
        R.close(); M.close()


@pytest.mark.parametrize("size,rows", [("narrow", 32), ("narrow", 130), ("narrow", 1024), ("base", 128)])
def test_f16_prompt_on_tensor_cores_opt_in_envelope(checkers, capi, zoo, size, rows):
    """F16 weights, opt-in (set_f16_tc_min_rows): the matmuls of 32+-row evals run on k_tcw_f16 (tcgen05 kind::f16, f32
    accumulation in TMEM, TMA-fed).  The tensor core adds the reference's products in its own order.  Measured envelope: the
    2-layer model meets the north star's f16 gate (logits within 1e-3 max-abs of the reference: 5e-4 .. 8e-4); on the 24-layer
    full-size synthetic model the same per-matmul noise (3e-7, tests/test_gpu_ops.py) is amplified by the fp16 re-roundings to
    1.9e-3 -- which is why the path is off by default.  Same argmax and top-5 everywhere, also for a greedy continuation on the KV
    cache that pass wrote; with the path off (the default) the same eval is bit-identical."""
    hp = {"narrow": gf.NARROW, "base": gf.BASE}[size]
    gate = {"narrow": 1e-3, "base": 3e-3}[size]
    p = zoo.path(size, "f16")
    toks = gf.synth_tokens(hp.n_positions, hp.n_vocab, seed=611)
    R = _fast_checker(checkers, p, n_batch=rows)
    M = capi.Model.load(p, max_batch=rows)
    assert M.eval_path(rows) == 0 and M.eval_path(1) == 3, capi.last_error()
    want = R.eval(toks[:rows], 0)
    got = M.eval(toks[:rows], 0)
    assert np.array_equal(_bits(got), _bits(want)), _diff(f"f16 {size} rows={rows} exact-order kernels (default)", got, want)
    M.set_f16_tc_min_rows(32)
    assert M.eval_path(rows) == 7 and M.eval_path(8) == 0 and M.eval_path(1) == 3, capi.last_error()
    got = M.eval(toks[:rows], 0)
    err = float(np.abs(got - want).max())
    print(f"f16 tcgen05 prompt {size} rows={rows}: max|d logits| = {err:.3e} (max|logit| {np.abs(want).max():.3f})")
    assert err <= gate, f"{size} rows={rows}: max|d|={err:.3e}"
    assert int(np.argmax(got)) == int(np.argmax(want)) and _top5(got) == _top5(want)
    if rows < hp.n_positions:
        tok = int(np.argmax(want))
        for i in range(12 if size == "narrow" else 4):
            lr = R.eval(np.array([tok], np.int32), rows + i)
            lm = M.eval(np.array([tok], np.int32), rows + i)
            assert float(np.abs(lm - lr).max()) <= gate and int(np.argmax(lm)) == int(np.argmax(lr)), (size, rows, i)
            tok = int(np.argmax(lr))
    R.close(); M.close()


def test_base_model_headline_greedy_over_whole_context(checkers, capi, zoo):
    """BASELINE configs[1] itself: 24 layers, vocabulary 42384, Q4_0, one token at a time from n_past 0 to 1023 on the
    persistent decode kernel (device-side greedy loop), against the reference's own greedy loop (biogpt.cpp:812-847 +
    top_k = 1 sampling): all 1024 ids identical; logits bit-identical where sampled (first steps, the 32-wide boundaries,
    the second K pass, the end of the context)"""
    hp = gf.BASE
    p = zoo.path("base", "q4_0")
    R = _fast_checker(checkers, p)
    M = capi.Model.load(p)
    assert M.eval_path(1) == 3 and M.decode_generation >= 4, capi.last_error()
    n = hp.n_positions
    ids, _ = M.decode_greedy(2, 0, n)
    probe = {0, 1, 2, 30, 31, 32, 33, 63, 64, 255, 256, 511, 512, 513, 767, 1021, 1022, 1023}
    want_ids, want_logits, tok = [], {}, 2
    for i in range(n):
        l = R.eval(np.array([tok], np.int32), i)
        if i in probe:
            want_logits[i] = (tok, l.copy())
        tok = int(np.argmax(l)); want_ids.append(tok)
    first_bad = next((i for i in range(n) if ids[i] != want_ids[i]), None)
    assert first_bad is None, f"greedy ids fork at step {first_bad}: {ids[first_bad]} vs {want_ids[first_bad]}"
    for i in sorted(probe):                      # the device cache holds this very sequence: re-evaluating position i is idempotent
        tok_i, wl = want_logits[i]
        got = M.eval(np.array([tok_i], np.int32), i)
        assert np.array_equal(_bits(got), _bits(wl)), _diff(f"base q4_0 n_past={i}", got, wl)
    R.close(); M.close()


@pytest.mark.parametrize("ftype", ["q4_1", "q5_0", "q5_1", "q8_0", "f16", "f32"])
def test_base_model_64_token_continuation(checkers, capi, zoo, ftype):
    """SURVEY 8(d) parity gate on the full-size model for every other format: an 8-token un-masked prompt batch, then a
    64-token greedy continuation -- ids identical, top-5 sets identical, and (stronger) logits bit-identical at every step"""
    hp = gf.BASE
    p = zoo.path("base", ftype)
    R = _fast_checker(checkers, p)
    M = capi.Model.load(p)
    prompt = gf.synth_tokens(8, hp.n_vocab, seed=17)
    want = R.eval(prompt, 0); got = M.eval(prompt, 0)
    assert np.array_equal(_bits(got), _bits(want)), _diff(f"{ftype} prompt", got, want)
    first = int(np.argmax(want))
    ids, _ = M.decode_greedy(first, 8, 64)
    tok, want_ids = first, []
    for i in range(64):
        l = R.eval(np.array([tok], np.int32), 8 + i)
        g = M.eval(np.array([tok], np.int32), 8 + i)
        assert _top5(g) == _top5(l), (ftype, i)
        assert np.array_equal(_bits(g), _bits(l)), _diff(f"{ftype} continuation step {i}", g, l)
        tok = int(np.argmax(l)); want_ids.append(tok)
    assert ids.tolist() == want_ids
    R.close(); M.close()


def test_eval_path_map(capi, zoo):
    """which schedule a (model, rows) pair takes: only quantised evals of 112+ rows leave the bit-exact kernels"""
    M = capi.Model.load(zoo.path("narrow", "q5_1"), max_batch=128)
    assert [M.eval_path(n) for n in (1, 2, 8, 9, 111, 127, 128, 1024)] == [3, 1, 1, 1, 1, 1, 4, 4]
    M.set_tc_min_rows(112)
    assert [M.eval_path(n) for n in (1, 8, 111, 112, 128)] == [3, 1, 1, 2, 2]
    M.set_tc_min_rows(0); M.set_tcx_min_rows(0)
    assert [M.eval_path(n) for n in (1, 8, 128, 1024)] == [3, 1, 1, 1]
    M.set_batch_path(2)                                                    # opt-in: persistent multi-row kernel for 2..8 rows
    assert [M.eval_path(n) for n in (2, 8, 9)] == [5, 5, 1]
    M.close()
    M = capi.Model.load(zoo.path("small", "q4_0"), max_batch=128)          # not BioGPT-base layer shapes: no skinny schedule
    assert [M.eval_path(n) for n in (1, 2, 32, 127, 128)] == [3, 0, 0, 0, 4]
    M.close()
    M = capi.Model.load(zoo.path("small", "f16"), max_batch=128)           # F16: exact-order SIMT kernels unless k_tcw_f16 is opted in
    assert [M.eval_path(n) for n in (1, 8, 32, 112)] == [3, 0, 0, 0]
    M.set_f16_tc_min_rows(32)
    assert [M.eval_path(n) for n in (1, 8, 31, 32, 112)] == [3, 0, 0, 7, 7]
    M.close()


def test_persistent_kernel_long_context(checkers, capi, zoo):
    """prompt in un-masked batches of 8 (per-op kernels), then persistent-kernel decode near the end
    of the context"""
    p = zoo.path("small", "q4_0")
    O = checkers.Oracle(p)
    M = capi.Model.load(p)
    toks = gf.synth_tokens(gf.SMALL.n_positions, gf.SMALL.n_vocab, seed=4)
    pos = 0
    while pos < 240:
        want = O.eval(toks[pos:pos + 8], pos); got = M.eval(toks[pos:pos + 8], pos)
        assert np.array_equal(_bits(got), _bits(want)), pos
        pos += 8
    for i in range(240, gf.SMALL.n_positions):
        want = O.eval(toks[i:i + 1], i); got = M.eval(toks[i:i + 1], i)
        assert np.array_equal(_bits(got), _bits(want)), _diff(f"n_past={i}", got, want)
    O.close(); M.close()


def test_errors_are_reported_not_swallowed(capi, zoo):
    M = capi.Model.load(zoo.path("tiny", "q4_0"))
    with pytest.raises(capi.BgptError):
        M.eval(np.zeros(4, np.int32), gf.TINY.n_positions - 2)   # runs past n_positions
    with pytest.raises(capi.BgptError):
        M.eval_streams(np.zeros(3, np.int32), 0)                 # streams not allocated
    M.close()

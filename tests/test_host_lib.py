"""The C++ host library (biogpt.cpp_b200/host): the reference's biogpt.h API on top of the C ABI.
CPU tier: text rules, BPE, sampler draw, the quantize front end (the reference's UNMODIFIED
quantize.cpp linked against our library) against the committed reference hashes.
GPU tier: biogpt_model_load + biogpt_eval through the C++ API against the oracle."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np
import pytest

from conftest import FTYPES, ROOT, gf

HOST = os.path.join(ROOT, "biogpt.cpp_b200", "host")
LIB = os.path.join(HOST, "libbiogpt_b200.so")


@pytest.fixture(scope="module")
def host():
    if not os.path.exists(LIB):
        subprocess.run(["make", "-C", ROOT, "product"], check=True, stdout=subprocess.DEVNULL)
    L = C.CDLL(LIB)
    L.bgpt_host_open.restype = C.c_void_p
    L.bgpt_host_open.argtypes = [C.c_char_p, C.c_int]
    L.bgpt_host_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.bgpt_host_n_vocab.argtypes = [C.c_void_p]
    L.bgpt_host_tokenize.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int]
    L.bgpt_host_close.argtypes = [C.c_void_p]
    L.bgpt_host_sample.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_uint32]
    L.bgpt_host_eval_sample.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_uint32]
    return L


def _tok(host, s):
    buf = C.create_string_buffer(1 << 16)
    n = host.bgpt_host_moses_tokenize(s.encode(), buf, len(buf))
    assert n >= 0
    return buf.value.decode().split("\n") if n else []


def test_moses_unit_strings(host):
    """the three strings of the reference's own unit test, mosestokenizer.cpp:491-497"""
    assert _tok(host, "Hello World!") == ["Hello", "World", "!"]
    assert _tok(host, "This ain't funny. It's actually hillarious, yet double Ls. | [] < > [ ] & You're gonna shake it off? Don't?") == [
        "This", "ain", "&apos;t", "funny", ".", "It", "&apos;s", "actually", "hillarious", ",", "yet", "double", "Ls", ".",
        "&#124;", "&#91;", "&#93;", "&lt;", "&gt;", "&#91;", "&#93;", "&amp;", "You", "&apos;re", "gonna", "shake", "it", "off", "?",
        "Don", "&apos;t", "?"]
    assert _tok(host, "this is a webpage https://stackoverflow.com/questions/6181381/how-to-print-variables-in-perl that kicks ass") == [
        "this", "is", "a", "webpage", "https", ":", "/", "/", "stackoverflow.com", "/", "questions", "/", "6181381", "/",
        "how", "@-@", "to", "@-@", "print", "@-@", "variables", "@-@", "in", "@-@", "perl", "that", "kicks", "ass"]


def test_moses_edge_cases(host):
    assert _tok(host, "") == []
    assert _tok(host, "   \t\n ") == []
    assert _tok(host, "Dr. Smith paid 1,000 dollars, e.g. in 1990's.") == [
        "Dr.", "Smith", "paid", "1,000", "dollars", ",", "e.g.", "in", "1990", "&apos;s", "."]
    buf = C.create_string_buffer(4096)
    host.bgpt_host_moses_detokenize("Hello\nWorld\n!\nIt\n&apos;s\nhow\n@-@\nto\n(\nfine\n)\n.".encode(), buf, len(buf))
    # AS BUILT the reference's detokenizer neither unescapes XML entities (it drops the result of its regex_replace,
    # mosestokenizer.cpp:386-390) nor attaches closing punctuation (its rule only matches one literal token, :419) and turns
    # " @-@" into "-" keeping the space after it; tests/test_text.py pins 327 strings against the reference's output
    assert buf.value.decode() == "Hello World ! It &apos;s how- to (fine ) ."


def test_bpe_lowest_rank_merges_first(host):
    buf = C.create_string_buffer(1024)
    merges = "l o\nlo w</w>\ne r</w>\nn e\nne w\nnew er</w>"
    host.bgpt_host_bpe(b"lower", merges.encode(), buf, len(buf))
    assert buf.value.decode() == "lo w e r</w>".replace("w e r</w>", "w er</w>")
    host.bgpt_host_bpe(b"newer", merges.encode(), buf, len(buf))
    assert buf.value.decode() == "newer</w>"
    host.bgpt_host_bpe(b"a", merges.encode(), buf, len(buf))
    assert buf.value.decode() == "a</w>"


def test_sampler_draw_identical_to_reference(host, checkers, zoo):
    """same libstdc++ mt19937 + discrete_distribution draw as biogpt_sample_top_k_top_p (biogpt.cpp:908-980)"""
    if not checkers.have_ref():
        pytest.skip("reference build not available")
    R = checkers.Ref(zoo.path("tiny", "f32"))
    rng = np.random.default_rng(0)
    for seed in range(20):
        logits = (rng.standard_normal(R.n_vocab) * 3).astype(np.float32)
        for top_k, top_p, temp in ((1, 1.0, 1.0), (40, 0.9, 0.9), (5, 0.5, 1.3), (R.n_vocab, 1.0, 0.7)):
            want = R.sample(logits, top_k, top_p, temp, seed)
            got = host.bgpt_host_sample(logits.ctypes.data, R.n_vocab, top_k, top_p, temp, seed)
            assert got == want, (seed, top_k, top_p, temp)
    R.close()


def test_sampler_fast_selection_equals_reference_on_ties_and_edges(host, checkers, zoo):
    """biogpt_sample_top_k_top_p selects the top_k + 1 largest logits in one pass instead of sorting the vocabulary; wherever
    std::partial_sort's outcome is not determined by the values alone (equal logits inside or at the edge of the top_k), for NaNs,
    for temperatures that collapse neighbours, and for top_k beyond the fast path, the reference's own code must decide: the drawn
    id equals the reference's for every case (biogpt.cpp:908-980)"""
    if not checkers.have_ref():
        pytest.skip("reference build not available")
    R = checkers.Ref(zoo.path("tiny", "f32"))
    n = R.n_vocab
    rng = np.random.default_rng(5)
    cases = []
    base = (rng.standard_normal(n) * 2).astype(np.float32)
    order = np.argsort(-base)
    x = base.copy(); x[order[3]] = x[order[2]]; cases.append(("tie inside the top", x))
    x = base.copy(); x[order[40]] = x[order[39]]; cases.append(("tie at the edge of top 40", x))
    x = base.copy(); x[order[41]] = x[order[40]]; cases.append(("tie just outside", x))
    x = base.copy(); x[order[5]] = x[order[4]]; x[order[6]] = x[order[4]]; cases.append(("triple", x))
    cases.append(("all equal", np.zeros(n, np.float32)))
    cases.append(("plateau of the maximum", np.where(np.arange(n) % 3 == 0, np.float32(1.5), base.clip(max=1.0)).astype(np.float32)))
    x = base.copy(); x[order[0]] = np.inf; cases.append(("+inf", x))
    x = base.copy(); x[order[10]] = -np.inf; x[7] = -np.inf; cases.append(("-inf", x))
    cases.append(("descending", np.linspace(5, -5, n).astype(np.float32)))
    cases.append(("ascending", np.linspace(-5, 5, n).astype(np.float32)))
    cases.append(("tiny values", (base * 1e-38).astype(np.float32)))
    for name, logits in cases:
        for top_k, top_p, temp in ((1, 1.0, 1.0), (5, 0.5, 1.3), (40, 0.9, 0.9), (40, 1.0, 1e300), (40, 0.9, 1e-300), (127, 0.95, 0.7), (n - 1, 1.0, 0.7), (n, 1.0, 0.7)):
            for seed in (0, 1, 2):
                want = R.sample(logits, top_k, top_p, temp, seed)
                got = host.bgpt_host_sample(logits.ctypes.data, n, top_k, top_p, temp, seed)
                assert got == want, (name, top_k, top_p, temp, seed)
    R.close()


def test_sampler_fast_selection_full_vocabulary(host):
    """the one-pass selection at the real vocabulary size (42384: chunk maxima give a lower bound, chunks below it are skipped) against
    the reference's partial_sort over all (logit / temp, id) pairs as written -- which test_sampler_* pin to the reference binary at the
    tiny model's size: same id for random rows, rows with ties inside / at the edge of the top_k, plateaus, infinities and NaNs"""
    host.bgpt_host_sample_n.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_uint32, C.c_int]
    n = 42384
    rng = np.random.default_rng(11)
    rows = [(rng.standard_normal(n) * s).astype(np.float32) for s in (1.0, 3.0, 1e-3)]
    base = rows[1]
    order = np.argsort(-base)
    x = base.copy(); x[order[7]] = x[order[6]]; rows.append(x)
    x = base.copy(); x[order[40]] = x[order[39]]; rows.append(x)
    x = base.copy(); x[order[41]] = x[order[40]]; rows.append(x)
    rows.append(np.where(np.arange(n) % 1000 == 0, np.float32(9.0), base.clip(max=8.0)).astype(np.float32))      # a plateau of 43 maxima
    x = base.copy(); x[order[0]] = np.inf; x[order[50]] = -np.inf; rows.append(x)
    x = base.copy(); x[12345] = np.nan; rows.append(x)
    rows.append(np.sort(base)); rows.append(np.sort(base)[::-1].copy())
    rows.append(np.zeros(n, np.float32))
    for r, logits in enumerate(rows):
        for top_k, top_p, temp in ((1, 1.0, 1.0), (40, 0.9, 0.8), (40, 1.0, 1.0), (100, 0.95, 1.2), (127, 0.5, 0.7), (300, 0.9, 1.0), (600, 0.9, 1.0)):
            for seed in (0, 1):
                want = host.bgpt_host_sample_n(logits.ctypes.data, n, top_k, top_p, temp, seed, 1)
                got = host.bgpt_host_sample_n(logits.ctypes.data, n, top_k, top_p, temp, seed, 0)
                assert got == want, (r, top_k, top_p, temp, seed)


@pytest.mark.parametrize("ftype", ["q4_0", "q4_1", "q5_0", "q5_1", "q8_0"])
def test_reference_quantize_frontend_on_our_library(zoo, model_dir, ftype):
    """the reference's UNMODIFIED examples/quantize/quantize.cpp, linked against libbiogpt_b200.so,
    must write byte-identical files to the reference's own quantize tool (hash in tests/golden)"""
    exe = os.path.join(HOST, "_build", "quantize")
    if not os.path.exists(exe):
        pytest.skip("front ends are only linked where /root/reference exists")
    gold = np.load(os.path.join(ROOT, "tests", "golden", "quantize_ref.npz"))
    out = os.path.join(model_dir, f"tiny-{ftype}-ourq.bin")
    tname = {"q4_0": "2", "q4_1": "3", "q5_0": "8", "q5_1": "9", "q8_0": "7"}[ftype]
    r = subprocess.run([exe, "-f", zoo.path("tiny", "f32"), "-o", out, "-t", tname], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-500:]
    assert hashlib.sha256(open(out, "rb").read()).digest() == gold[f"sha_{ftype}"].tobytes()


def test_main_frontend_links():
    exe = os.path.join(HOST, "_build", "biogpt")
    if not os.path.exists(exe):
        pytest.skip("front ends are only linked where /root/reference exists")
    r = subprocess.run([exe, "-h"], capture_output=True, text=True)
    assert "usage:" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("ftype", ["q4_0", "f16", "q5_1"])
def test_cpp_api_eval_matches_oracle(host, checkers, zoo, ftype):
    """biogpt_model_load -> measure pass -> biogpt_eval (the calls main.cpp makes) == oracle bits"""
    p = zoo.path("small", ftype)
    h = host.bgpt_host_open(p.encode(), 8)
    assert h
    O = checkers.Oracle(p)
    toks = gf.synth_tokens(20, gf.SMALL.n_vocab, seed=12)
    out = np.zeros(host.bgpt_host_n_vocab(h), np.float32)
    pos = 0
    for n in (8, 4, 1, 1, 1, 5):
        t = np.ascontiguousarray(toks[pos:pos + n])
        assert host.bgpt_host_eval(h, t.ctypes.data, n, pos, out.ctypes.data) == 0
        want = O.eval(t, pos)
        assert np.array_equal(out.view(np.uint32), want.view(np.uint32)), (ftype, pos, n)
        pos += n
    ids = np.zeros(16, np.int32)
    n = host.bgpt_host_tokenize(h, b"a b c", ids.ctypes.data, 16)
    assert ids[:n].tolist() == [2, 4, 5, 6]          # "a b c" -> 2 4 5 6
    host.bgpt_host_close(h); O.close()


@pytest.mark.gpu
@pytest.mark.parametrize("size,ftype", [("small", "q4_0"), ("narrow", "q5_1"), ("small", "f16")])
def test_device_topk_sampler_draws_the_reference_id(host, checkers, zoo, size, ftype):
    """biogpt_eval_sample (top_k selection on the device, csrc/bgpt_topk.cuh; softmax / top-p / mt19937 draw on the host) returns the
    id that the reference's biogpt_sample_top_k_top_p (biogpt.cpp:908-980) draws from the full logits: 20 seeds x the 4 settings of
    test_sampler_draw_identical_to_reference, on a prompt batch and on decode steps"""
    if not checkers.have_ref():
        pytest.skip("reference build not available")
    hp = {"small": gf.SMALL, "narrow": gf.NARROW}[size]
    p = zoo.path(size, ftype)
    h = host.bgpt_host_open(p.encode(), 8)
    assert h
    R = checkers.Ref(p)
    toks = gf.synth_tokens(16, hp.n_vocab, seed=5)
    logits = np.zeros(hp.n_vocab, np.float32)
    evals = [(toks[:8], 0)] + [(toks[8 + i:9 + i], 8 + i) for i in range(4)]
    for t, n_past in evals:
        t = np.ascontiguousarray(t)
        assert host.bgpt_host_eval(h, t.ctypes.data, len(t), n_past, logits.ctypes.data) == 0
        for seed in range(20):
            for top_k, top_p, temp in ((1, 1.0, 1.0), (40, 0.9, 0.9), (5, 0.5, 1.3), (128, 1.0, 0.7)):
                want = R.sample(logits, top_k, top_p, temp, seed)
                got = host.bgpt_host_eval_sample(h, t.ctypes.data, len(t), n_past, top_k, top_p, temp, seed)
                assert got == want, (n_past, seed, top_k, top_p, temp)
    host.bgpt_host_close(h); R.close()


def test_loader_rejects_corrupt_files_without_throwing(host, zoo, model_dir):
    """sizes read from the file are validated before use: a truncated file, an absurd vocabulary string length and a negative
    tensor dimension all make biogpt_model_load return false (no exception, no multi-GB allocation)"""
    import struct
    good = open(zoo.path("tiny", "q4_0"), "rb").read()
    cases = {
        "truncated_vocab": good[:200],
        "huge_vocab_len": good[:36] + struct.pack("<I", 0x7FFFFFF0) + good[40:],
        "truncated_tensor": good[:len(good) - 1000],
    }
    # a negative ne[0] in the first tensor header: find the first tensor by its name
    at = good.find(b"biogpt.embed_tokens.weight")
    hdr = at - 8 - 12                                   # n_dims, name_len, ttype, ne[0], ne[1] precede the name
    cases["negative_dim"] = good[:hdr + 12] + struct.pack("<i", -64) + good[hdr + 16:]
    for name, blob in cases.items():
        p = os.path.join(model_dir, f"corrupt-{name}.bin")
        open(p, "wb").write(blob)
        assert not host.bgpt_host_open(p.encode(), 8), name


@pytest.mark.gpu
@pytest.mark.parametrize("ftype", ["q4_0", "f16"])
def test_reference_main_binary_end_to_end(checkers, zoo, ftype):
    """the reference's UNMODIFIED examples/main/main.cpp, linked against libbiogpt_b200.so (host/Makefile), run end to end on the GPU:
    `biogpt -m model -p "a b c" -n 12 --top_k 1` -- tokenizer, prompt batch, greedy sampling, detokenizer -- must print the ids the
    oracle's greedy loop produces (synthetic vocabulary: id 30 + i decodes to "tok{i}")"""
    import re
    exe = os.path.join(HOST, "_build", "biogpt")
    if not os.path.exists(exe):
        pytest.skip("front ends are only linked where /root/reference exists (the built binary travels to the GPU box)")
    p = zoo.path("small", ftype)
    r = subprocess.run([exe, "-m", p, "-p", "a b c", "-n", "12", "--top_k", "1", "-s", "3"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-800:]
    assert "number of tokens in prompt = 4, first 8 tokens: 2 4 5 6" in r.stdout, r.stdout[-800:]
    body = r.stdout.split("first 8 tokens:")[1].split("load time")[0]
    got = [30 + int(x) for x in re.findall(r"tok(\d+)", body)]
    O = checkers.Oracle(p)
    l = O.eval(np.array([2, 4, 5, 6], np.int32), 0)
    want = []
    for i in range(12):
        tok = int(np.argmax(l)); want.append(tok)
        l = O.eval(np.array([tok], np.int32), 4 + i)
    O.close()
    want_printed = [t for t in want if t >= 30]          # ids below 30 are "<s>", "a".."z" etc.
    assert got == want_printed, (got, want)

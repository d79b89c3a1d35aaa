"""Host text code (host/text.cpp) against the AS-BUILT reference tokenizer (/root/reference/mosestokenizer.cpp): token lists and
detokenised strings for 335 strings -- abbreviations, runs of dots, numbers with commas, apostrophes, hyphens, brackets, non-ASCII
punctuation -- committed as tests/golden/moses_golden.json (tests/golden/make_moses_golden.py), and live against the reference
library where /root/reference exists.  CPU tier."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from conftest import ROOT

LIB = os.path.join(ROOT, "biogpt.cpp_b200", "host", "libbiogpt_b200.so")
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libmoses_ref.so")
REF_CWD = "/root/reference/examples"


@pytest.fixture(scope="module")
def host():
    L = C.CDLL(LIB)
    L.bgpt_host_moses_tokenize.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    L.bgpt_host_moses_detokenize.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    return L


def _tok(host, s: str):
    buf = C.create_string_buffer(1 << 16)
    n = host.bgpt_host_moses_tokenize(s.encode("utf-8", errors="surrogateescape"), buf, len(buf))
    assert n >= 0
    return buf.value.decode("utf-8", errors="surrogateescape").split("\n") if n else []


def _detok(host, toks):
    buf = C.create_string_buffer(1 << 16)
    n = host.bgpt_host_moses_detokenize("\n".join(toks).encode("utf-8", errors="surrogateescape"), buf, len(buf))
    assert n >= 0
    return buf.value.decode("utf-8", errors="surrogateescape")


def _cases():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "moses_golden.json")))["cases"]


def test_tokenizer_matches_reference_golden(host):
    bad = []
    for c in _cases():
        if c.get("reference_throws"):
            _tok(host, c["text"])                      # the reference aborts on these; ours must simply not
            continue
        got = _tok(host, c["text"])
        if got != c["tokens"]:
            bad.append((c["text"], got, c["tokens"]))
    assert not bad, f"{len(bad)} of {len(_cases())} strings tokenise differently, e.g. {bad[:3]}"


def test_detokenizer_matches_reference_golden(host):
    bad = []
    for c in _cases():
        if c.get("reference_throws"):
            continue
        got = _detok(host, c["tokens"])
        if got != c["detok"]:
            bad.append((c["tokens"], got, c["detok"]))
    assert not bad, f"{len(bad)} detokenise differently, e.g. {bad[:3]}"


def test_builtin_byte_classes_equal_the_reference_data_files(host):
    """host/text.cpp carries the five byte sets built in (used when ../data is not reachable from the cwd, as here): they must be
    exactly the sets of bytes occurring in the reference's data/perluniprops files"""
    base = "/root/reference/data/perluniprops"
    if not os.path.isdir(base):
        pytest.skip("reference data not available")
    host.bgpt_host_text_class_mask.argtypes = [C.c_int, C.c_void_p]
    for which, cat in enumerate(["IsAlnum", "IsAlpha", "IsLower", "IsN", "IsSc"]):
        m = np.zeros(8, np.uint32)
        host.bgpt_host_text_class_mask(which, m.ctypes.data)
        mine = {b for b in range(256) if (int(m[b >> 5]) >> (b & 31)) & 1}
        want = set(open(os.path.join(base, cat + ".txt"), "rb").read())
        assert mine == want, (cat, sorted(mine ^ want))


def test_tokenizer_matches_live_reference(host, tmp_path):
    """the same corpus plus fresh random strings against the reference library itself, run in its own process (builder container only)"""
    if not (os.path.exists(REF_LIB) and os.path.isdir(REF_CWD)):
        pytest.skip("reference tokenizer not built here")
    import random
    import subprocess
    import sys
    rng = random.Random(7)
    alphabet = list("abcXYZ019 .,'-()[]\"?!;:&|<>%$") + ["...", " Dr.", " al.", " e.g.", "\u00e9", "\u03b1", "\u201c", "\u201d", "\u2026", " 1,000", " it's", "\t", "  "]
    texts = [c["text"] for c in _cases()] + ["".join(rng.choice(alphabet) for _ in range(rng.randint(1, 40))) for _ in range(400)]
    fin, fout = str(tmp_path / "in.json"), str(tmp_path / "out.json")
    json.dump(texts, open(fin, "w"))
    subprocess.run([sys.executable, os.path.join(ROOT, "tests", "golden", "make_moses_golden.py"), "--tokenize", fin, fout], check=True)
    wants = json.load(open(fout))
    bad = [(t, _tok(host, t), w) for t, w in zip(texts, wants) if w is not None and _tok(host, t) != w]
    assert not bad, f"{len(bad)} of {len(texts)} differ, e.g. {bad[:3]}"

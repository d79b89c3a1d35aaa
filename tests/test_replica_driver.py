"""The C++ multi-GPU replica driver (include/bgpt_replicas.h, host/replicas.cpp): one host thread + one engine per device,
stream s on device s % G, lock-step streams per device.  CPU tier: the symbols exist and the driver fails loudly without a GPU.
GPU tier: every stream of the driver equals the single-stream engine bit for bit / id for id, on however many GPUs the box has."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, gf, have_gpu

LIB = os.path.join(ROOT, "biogpt.cpp_b200", "host", "libbiogpt_b200.so")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def rep():
    L = C.CDLL(LIB)
    L.bgpt_replicas_open.restype = C.c_void_p
    L.bgpt_replicas_open.argtypes = [C.c_char_p, C.c_int, C.c_int]
    L.bgpt_replicas_close.argtypes = [C.c_void_p]
    for f in ("devices", "streams", "n_vocab"):
        getattr(L, "bgpt_replicas_" + f).argtypes = [C.c_void_p]
    L.bgpt_replicas_device_of.argtypes = [C.c_void_p, C.c_int]
    L.bgpt_replicas_eval.argtypes = [C.c_void_p, _i32p, C.c_int, _f32p]
    L.bgpt_replicas_decode_greedy.argtypes = [C.c_void_p, _i32p, C.c_int, C.c_int, _i32p, C.c_void_p]
    L.bgpt_replicas_eval_topk.argtypes = [C.c_void_p, _i32p, C.c_int, C.c_int, _f32p, _i32p, _i32p, _i32p, C.c_void_p]
    L.bgpt_replicas_last_error.restype = C.c_char_p
    return L


def test_replica_header_symbols_are_exported(rep):
    src = open(os.path.join(ROOT, "include", "bgpt_replicas.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = sorted(set(re.findall(r"\b(bgpt_replicas_[a-z0-9_]+)\s*\(", src)))
    assert len(names) >= 9
    for n in names:
        assert hasattr(rep, n), n


@pytest.mark.skipif(have_gpu(), reason="checks the behaviour WITHOUT a device")
def test_replica_driver_fails_loudly_without_gpu(rep, zoo):
    assert not rep.bgpt_replicas_open(zoo.path("tiny", "q4_0").encode(), 0, 4)
    assert b"no CUDA device" in rep.bgpt_replicas_last_error()


@pytest.mark.gpu
@pytest.mark.parametrize("size,ftype,S", [("narrow", "q5_1", 5), ("small", "f16", 3), ("narrow", "q4_0", 16)])
def test_replica_streams_equal_single_stream(rep, capi, zoo, size, ftype, S):
    hp = {"narrow": gf.NARROW, "small": gf.SMALL}[size]
    p = zoo.path(size, ftype)
    steps = 12
    seqs = [gf.synth_tokens(steps, hp.n_vocab, seed=500 + s) for s in range(S)]
    M = capi.Model.load(p)
    single = [np.stack([M.eval(seqs[s][i:i + 1], i) for i in range(steps)]) for s in range(S)]
    greedy = [M.decode_greedy(int(seqs[s][0]), 0, steps)[0] for s in range(S)]
    M.close()
    r = rep.bgpt_replicas_open(p.encode(), 0, S)
    assert r, rep.bgpt_replicas_last_error()
    G = rep.bgpt_replicas_devices(r)
    assert G >= 1 and rep.bgpt_replicas_streams(r) == S and rep.bgpt_replicas_n_vocab(r) == hp.n_vocab
    assert [rep.bgpt_replicas_device_of(r, s) for s in range(S)] == [s % G for s in range(S)]
    out = np.zeros((S, hp.n_vocab), np.float32)
    for i in range(steps):
        toks = np.array([seqs[s][i] for s in range(S)], np.int32)
        assert rep.bgpt_replicas_eval(r, toks, i, out) == 0, rep.bgpt_replicas_last_error()
        for s in range(S):
            assert np.array_equal(out[s].view(np.uint32), single[s][i].view(np.uint32)), (ftype, s, i)
    # the sampler's form of the same step: per stream the 40 best (logit, id) pairs
    K = 40
    vals = np.zeros((S, K), np.float32); tids = np.zeros((S, K), np.int32); n_out = np.zeros(S, np.int32); exact = np.zeros(S, np.int32)
    fb = np.zeros((S, hp.n_vocab), np.float32)
    i = steps - 1
    toks = np.array([seqs[s][i] for s in range(S)], np.int32)
    assert rep.bgpt_replicas_eval_topk(r, toks, i, K, vals.reshape(-1), tids.reshape(-1), n_out, exact, fb.ctypes.data) == 0, rep.bgpt_replicas_last_error()
    for s in range(S):
        order = np.argsort(-single[s][i], kind="stable")[:K]
        if exact[s]:
            assert n_out[s] == K and tids[s].tolist() == order.tolist() and np.array_equal(vals[s].view(np.uint32), single[s][i][order].view(np.uint32)), (ftype, s)
        else:
            assert np.array_equal(fb[s].view(np.uint32), single[s][i].view(np.uint32)), (ftype, s)
    ids = np.zeros((steps, S), np.int32)
    ms = np.zeros(G, np.float32)
    first = np.array([seqs[s][0] for s in range(S)], np.int32)
    assert rep.bgpt_replicas_decode_greedy(r, first, 0, steps, ids.reshape(-1), ms.ctypes.data) == 0, rep.bgpt_replicas_last_error()
    for s in range(S):
        assert ids[:, s].tolist() == greedy[s].tolist(), (ftype, s)
    assert (ms > 0).all()
    rep.bgpt_replicas_close(r)

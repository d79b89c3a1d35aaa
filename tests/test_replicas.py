"""N > 1 host logic on CPU: world_size-2 `gloo` process group.

The data path has no collective (replicas only); what is checked here is the
plumbing around it: stream -> rank assignment, barrier, MAX-over-ranks time / SUM tokens, and
that a job sharded over 2 ranks yields, stream by stream, the ids of the unsharded job.  The
"engine" of each rank in this CPU test is the oracle on the tiny model (tests may use it as a
stand-in; on the GPU box bench.py runs the CUDA engine through the same functions).
"""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, gf

import importlib

replicas = importlib.import_module("biogpt_cpp_b200.replicas")


def test_assign_streams_partition():
    for n in (0, 1, 7, 8, 64):
        for w in (1, 2, 3, 8):
            owned = [replicas.assign_streams(n, w, r) for r in range(w)]
            flat = sorted(s for o in owned for s in o)
            assert flat == list(range(n))
            assert [len(o) for o in owned] == replicas.streams_per_rank(n, w)
            assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1
            for r, o in enumerate(owned):
                assert all(s % w == r for s in o)
    with pytest.raises(ValueError):
        replicas.assign_streams(4, 2, 2)


def test_single_process_passthrough():
    assert replicas.reduce_job(None, 10, 2.0) == (10.0, 2.0)
    assert replicas.job_tokens_per_s(None, 10, 2.0) == 5000.0
    assert replicas.gather_ids(None, [[1, 2]], 1, 1, 0) == [[1, 2]]
    replicas.barrier(None)


def _greedy(path, first, steps):
    import ref
    O = ref.Oracle(path)
    ids, tok = [], np.array([first], dtype=np.int32)
    for p in range(steps):
        tok[0] = int(np.argmax(O.eval(tok, p)))
        ids.append(int(tok[0]))
    O.close()
    return ids


def _worker(rank, world, port, path, n_streams, steps, q):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from _bootstrap import load_pkg
    load_pkg()
    rep = importlib.import_module("biogpt_cpp_b200.replicas")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = rep.assign_streams(n_streams, world, rank)
        rep.barrier(dist)
        ids = [_greedy(path, 4 + s, steps) for s in mine]
        ms_local = 10.0 * (rank + 1)                   # pretend device times: rank 1 is the slow one
        rep.barrier(dist)
        tokens, ms = rep.reduce_job(dist, len(mine) * steps, ms_local)
        tps = rep.job_tokens_per_s(dist, len(mine) * steps, ms_local)
        allids = rep.gather_ids(dist, ids, n_streams, world, rank)
        q.put((rank, tokens, ms, tps, allids))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


@pytest.mark.timeout(300)
def test_gloo_world2_matches_single_process(zoo, checkers):
    import torch.multiprocessing as mp
    path = zoo.path("tiny", "q4_0")
    n_streams, steps, world = 5, 6, 2
    want = [_greedy(path, 4 + s, steps) for s in range(n_streams)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, path, n_streams, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort()
    for rank, tokens, ms, tps, allids in got:
        assert tokens == n_streams * steps            # SUM over ranks
        assert ms == 20.0                             # MAX over ranks
        assert tps == pytest.approx(n_streams * steps / 0.020)
        if rank == 0:
            assert allids == want
        else:
            assert allids is None

"""Replica sharding of independent prompt streams over GPUs.

One decode sequence is a strict dependency chain and the whole model fits one GPU, so the path
does not shard: every rank (one process per GPU) holds a full weight copy and owns the streams
`s` with `s % world == rank` (stream s -> GPU s mod G), each with its own KV cache.  There is no
collective on the data path.  The only cross-rank traffic is bookkeeping for the measurement:
a barrier around the timed region and a MAX (time) / SUM (tokens) reduction of two scalars.

Nothing here touches CUDA directly: the reductions run on whatever backend the process group was
created with (NCCL on the GPU box, gloo in the CPU tests of tests/test_replicas.py).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple


def assign_streams(n_streams: int, world: int, rank: int) -> List[int]:
    """global stream ids owned by `rank`: s % world == rank, ascending"""
    if world < 1 or not (0 <= rank < world) or n_streams < 0:
        raise ValueError(f"assign_streams: n_streams={n_streams} world={world} rank={rank}")
    return list(range(rank, n_streams, world))


def streams_per_rank(n_streams: int, world: int) -> List[int]:
    """how many streams each rank owns (differs by at most one)"""
    return [len(range(r, n_streams, world)) for r in range(world)]


def _tensor(dist, vals: Sequence[float]):
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    return torch.tensor(list(vals), dtype=torch.float64, device=dev)


def barrier(dist) -> None:
    """device-complete barrier (None = single process)"""
    if dist is None:
        return
    if dist.get_backend() == "nccl":
        import torch
        torch.cuda.synchronize()
    dist.barrier()


def reduce_job(dist, tokens_local: float, ms_local: float) -> Tuple[float, float]:
    """whole-job (tokens, milliseconds): tokens summed over ranks, time = MAX over ranks.
    Every rank gets the same pair."""
    if dist is None:
        return float(tokens_local), float(ms_local)
    t = _tensor(dist, [ms_local])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n = _tensor(dist, [tokens_local])
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    return float(n.item()), float(t.item())


def job_tokens_per_s(dist, tokens_local: float, ms_local: float) -> float:
    n, ms = reduce_job(dist, tokens_local, ms_local)
    return n / (ms / 1e3) if ms > 0 else 0.0


def gather_ids(dist, ids_local: Sequence[Sequence[int]], n_streams: int, world: int, rank: int) -> Optional[List[List[int]]]:
    """rank 0 receives every stream's sampled ids in global stream order (other ranks: None).
    Used by the tests to check that a sharded job equals the single-process job stream by stream."""
    mine = assign_streams(n_streams, world, rank)
    if len(mine) != len(ids_local):
        raise ValueError("gather_ids: one id list per owned stream")
    if dist is None:
        return [list(x) for x in ids_local]
    payload = {s: list(x) for s, x in zip(mine, ids_local)}
    out = [None] * world if rank == 0 else None
    dist.gather_object(payload, out, dst=0)
    if rank != 0:
        return None
    merged = {}
    for part in out:
        merged.update(part)
    return [merged[s] for s in range(n_streams)]

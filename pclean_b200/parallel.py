"""Multi-process host glue of the row-sharded engine, used by
bench.py and exercised on CPU with world_size-2 gloo (tests/test_parallel_gloo.py).

Rows of the observation class have no incoming references (inference.jl:1-2): contiguous row
ranges per rank, latent tables replicated.  Per sweep the engine itself all-reduces the reference
counts over NCCL and all-gathers the rows proposed as new, which every replica then creates in
(rank, row) order so that all replicas assign the same slots; the helpers here are the host side
of that protocol: the shard ranges, the hand-over of the NCCL unique id through torch.distributed,
the replay order of gathered records, and the max-over-ranks reduction of timings."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def shard_range(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced, covering row range of `rank`."""
    return (n_rows * rank) // world, (n_rows * (rank + 1)) // world


def broadcast_bytes(payload: bytes, n: int, src: int = 0, device: str = "cpu") -> bytes:
    """Rank `src` hands `n` bytes (the 128-byte NCCL unique id) to every rank."""
    import torch
    import torch.distributed as dist
    t = torch.zeros(n, dtype=torch.uint8, device=device)
    if dist.get_rank() == src:
        t.copy_(torch.tensor(list(payload[:n]), dtype=torch.uint8))
    dist.broadcast(t, src)
    return bytes(t.cpu().tolist())


def attach_row_shard(engine, cls: int, n_rows: int, rank: int, world: int, device: str = "cuda") -> Tuple[int, int]:
    """Shard the observation class of `engine` over the ranks of the default process group: rank 0
    creates the NCCL unique id, everybody joins the engine's own communicator."""
    r0, r1 = shard_range(n_rows, rank, world)
    if world > 1:
        uid = broadcast_bytes(type(engine).nccl_unique_id() if rank == 0 else b"", 128, 0, device)
        engine.set_row_shard(cls, r0, r1)
        engine.nccl_init(uid, rank, world)
    return r0, r1


def replay_order(counts: Sequence[int], max_count: int) -> List[int]:
    """Index, into the all-gathered fixed-capacity record buffer [world][max_count], of every record in
    the order all replicas create the proposed rows: by rank, then by row (what engine.cu's
    exchange path computes before k_unpack_requests)."""
    return [rk * max_count + i for rk, c in enumerate(counts) for i in range(c)]


def gather_requests(local_rows: np.ndarray, local_records: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """All-gather variable-length (row id, record) requests with a count header and return them in
    replay order — the host-side statement of the engine's new-row exchange (gloo on CPU)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    cnt = torch.tensor([len(local_rows)], dtype=torch.int64)
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, cnt)
    counts = [int(c) for c in counts]
    mx = max(1, max(counts))
    w = local_records.shape[1] if local_records.ndim == 2 else 0
    buf = torch.zeros((mx, w + 1), dtype=torch.int64)
    if len(local_rows):
        buf[:len(local_rows), 0] = torch.from_numpy(np.asarray(local_rows, dtype=np.int64))
        buf[:len(local_rows), 1:] = torch.from_numpy(np.asarray(local_records, dtype=np.int64))
    allb = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(allb, buf)
    flat = torch.cat(allb).numpy()
    order = replay_order(counts, mx)
    return flat[order, 0], flat[order, 1:]


def max_over_ranks(values: Sequence[float], device: str = "cpu") -> List[float]:
    """Timings are reported as the maximum over ranks."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]

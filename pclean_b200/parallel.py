"""Row sharding of the observation class across ranks.

Rows of the observation class have no incoming references (inference.jl:1-2), so given the
sweep's table snapshot every row move is independent: contiguous row ranges per rank, latent
tables replicated, and ONE all-reduce(sum) of the reference counts per sweep.  The engine
does that all-reduce itself over NCCL (pclean_nccl_init); the helpers here are the host-side
arithmetic, also exercised on CPU with the gloo backend (tests/test_parallel_gloo.py)."""
from __future__ import annotations

from typing import Tuple

import numpy as np


def shard_range(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced, covering row range of `rank`."""
    return (n_rows * rank) // world, (n_rows * (rank + 1)) // world


def local_reference_counts(assignment_slots: np.ndarray, capacity: int, begin: int, end: int) -> np.ndarray:
    """Histogram of the target slots referenced by rows [begin, end) — what k_count_assign
    computes on the device for one rank's shard."""
    return np.bincount(assignment_slots[begin:end], minlength=capacity).astype(np.int32)


def allreduce_counts(counts: np.ndarray) -> np.ndarray:
    """Sum the per-rank histograms (torch.distributed; gloo on CPU, NCCL on GPU)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(counts))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy()

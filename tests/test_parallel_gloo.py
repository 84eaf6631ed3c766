"""world_size-2 gloo test of the N>1 host logic: shard ranges cover the rows exactly and the
all-reduced per-shard reference counts equal the global histogram (the one collective of a sweep)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pclean_b200.parallel import allreduce_counts, local_reference_counts, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_rows, cap, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)
    slots = rng.integers(0, cap, size=n_rows)
    b, e = shard_range(n_rows, rank, world)
    total = allreduce_counts(local_reference_counts(slots, cap, b, e))
    if rank == 0:
        q.put((total.tolist(), np.bincount(slots, minlength=cap).tolist()))
    dist.destroy_process_group()


def test_shard_ranges_cover():
    for n in (0, 1, 7, 1000, 1_000_003):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1


def test_allreduce_counts_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 10007, 64, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, want = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == want

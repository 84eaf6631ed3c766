"""world_size-2 gloo tests of the N>1 host logic bench.py runs (pclean_b200/parallel.py): shard
ranges, the hand-over of the NCCL unique id, the (rank, row) replay order of the new-row exchange —
every replica must create the gathered rows in the same order, the one a single process would
use — and the max-over-ranks timing reduction."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from pclean_b200.parallel import broadcast_bytes, gather_requests, max_over_ranks, replay_order, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _requests(n_rows):
    """rows that ask for a new latent row (every 7th) with a 3-int record each"""
    rows = np.arange(0, n_rows, 7)
    return rows, np.stack([rows * 3 + 1, rows % 5, rows // 2], axis=1)


def _worker(rank, world, port, n_rows, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    uid = broadcast_bytes(bytes(range(128)) if rank == 0 else b"", 128)
    b, e = shard_range(n_rows, rank, world)
    rows, recs = _requests(n_rows)
    mine = (rows >= b) & (rows < e)
    got_rows, got_recs = gather_requests(rows[mine], recs[mine])
    t = max_over_ranks([1.0 + rank, 5.0 - rank])
    q.put((rank, uid == bytes(range(128)), got_rows.tolist(), got_recs.tolist(), t))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover():
    for n in (0, 1, 7, 1000, 1_000_003):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1


def test_replay_order_is_rank_then_row():
    assert replay_order([2, 0, 3], 4) == [0, 1, 8, 9, 10]


def test_exchange_protocol_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n_rows = 1003
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_rows, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rows, recs = _requests(n_rows)
    for rank, uid_ok, got_rows, got_recs, t in res:
        assert uid_ok
        assert got_rows == rows.tolist() and got_recs == recs.tolist()      # contiguous shards: (rank, row) order = row order
        assert t == [2.0, 5.0]

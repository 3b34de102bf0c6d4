"""CPU, world_size 2 over gloo: the N>1 host logic (stream sharding, weight-blob broadcast,
max-over-ranks timing).  The data path itself has no collective."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from whisperlivekit_b200.sharding import StickyPlacement, broadcast_blob, max_over_ranks, shard_streams


def test_shard_streams_balanced_and_disjoint():
    for n, w in [(512, 8), (7, 2), (3, 4), (64, 1)]:
        sh = shard_streams(n, w)
        assert sorted(x for s in sh for x in s) == list(range(n))
        assert max(map(len, sh)) - min(map(len, sh)) <= 1


def test_sticky_placement():
    p = StickyPlacement(4)
    ranks = [p.open(i) for i in range(10)]
    assert sorted(p.load) == [2, 2, 3, 3]
    p.close(0)
    assert p.open(100) == ranks[0]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    blob = torch.zeros(1 << 16, dtype=torch.uint8)
    if rank == 0:
        blob.copy_(torch.from_numpy(np.random.default_rng(0).integers(0, 255, 1 << 16, dtype=np.uint8)))
    broadcast_blob(blob, src=0)
    mine = shard_streams(9, world)[rank]
    t = max_over_ranks(1.0 + rank)
    q.put((rank, int(blob.sum()), mine, t))
    dist.destroy_process_group()


def test_two_rank_broadcast_and_sharding():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(30)
    assert res[0][1] == res[1][1] != 0
    assert res[0][2] == [0, 1, 2, 3, 4] and res[1][2] == [5, 6, 7, 8]
    assert res[0][3] == res[1][3] == 2.0

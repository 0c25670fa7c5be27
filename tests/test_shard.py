"""CPU: the N>1 host logic (pair sharding, max-over-ranks timing) with world_size 2 on the gloo backend."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vtm_b200.shard import gather_order, max_over_ranks, pairs_for_rank, sum_over_ranks


def test_pairs_partition():
    for n in (1, 7, 256):
        for world in (1, 2, 4, 8):
            parts = [pairs_for_rank(n, world, r) for r in range(world)]
            flat = sorted(p for part in parts for p in part)
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
            assert sorted(gather_order(n, world)) == list(range(n))
    with pytest.raises(ValueError):
        pairs_for_rank(4, 2, 2)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = pairs_for_rank(9, world, rank)
    # every rank "processes" its pairs; rank-dependent fake time
    t = max_over_ranks(10.0 + rank)
    total = sum_over_ranks(len(mine))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        torch.save({"t": t, "total": total, "gathered": gathered}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "r.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["t"] == 11.0            # max over ranks, not rank 0's own time
    assert r["total"] == 9.0
    assert sorted(r["gathered"][0] + r["gathered"][1]) == list(range(9))
    assert not set(r["gathered"][0]) & set(r["gathered"][1])

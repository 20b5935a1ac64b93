"""world_size-2 gloo test of the only multi-GPU logic the path has: image sharding + max/sum timing reduction."""
import os
import socket

import pytest

from heif_b200.sharding import reduce_timing, shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 48, 4096, 4097):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                seen += list(shard_range(n, r, world))
            assert seen == list(range(n))
            sizes = [len(shard_range(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_range(101, rank, world)
    ms, sums = reduce_timing(dist, torch.device("cpu"), 10.0 + rank, {"images": len(mine), "bins": 1000 * (rank + 1)})
    dist.barrier()
    q.put((rank, ms, sums, (mine.start, mine.stop)))
    dist.destroy_process_group()


def test_two_rank_reduction_over_gloo():
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ms, sums, _ in res:
        assert ms == 11.0                      # max over ranks
        assert sums == {"bins": 3000.0, "images": 101.0}
    assert res[0][3] == (0, 51) and res[1][3] == (51, 101)

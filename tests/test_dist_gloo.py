"""N>1 host logic on CPU: two gloo ranks shard the games, broadcast a weight blob and reduce timings."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    from minizero_b200 import dist as mzdist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    games = mzdist.shard_games(256, world, rank)
    blob = torch.arange(4096, dtype=torch.uint8) if rank == 0 else torch.zeros(4096, dtype=torch.uint8)
    mzdist.broadcast_blob(dist, blob, src=0)
    dims = mzdist.broadcast_object(dist, {"num_blocks": 6} if rank == 0 else None, src=0)
    mx = mzdist.max_over_ranks(dist, [10.0 + rank, 5.0 - rank])
    sm = mzdist.sum_over_ranks(dist, [float(len(games))])
    out[rank] = (games[:3], len(games), int(blob.sum()), dims, mx, sm)
    dist.destroy_process_group()


def test_two_ranks_shard_broadcast_reduce():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    expect_sum = int(torch.arange(4096, dtype=torch.uint8).sum())
    assert out[0][0] == [0, 2, 4] and out[1][0] == [1, 3, 5]
    assert out[0][1] + out[1][1] == 256
    for r in range(world):
        assert out[r][2] == expect_sum and out[r][3] == {"num_blocks": 6}
        assert out[r][4] == [11.0, 5.0] and out[r][5] == [256.0]

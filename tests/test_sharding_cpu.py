"""N > 1 host logic on CPU: world_size-2 gloo processes exercise the sharding and the max-over-ranks
reduction that bench.py uses around its timed region (the data path itself has no collective)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from coin_b200 import sharding  # importing the package needs libcoinops.so (built by build()); no GPU needed


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.shard_images(7, rank, world)
    ms = [10.0 + rank, 3.0 - rank]                       # rank-dependent "measured times"
    mx = sharding.max_over_ranks(ms)
    dist.barrier()
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        out.put((gathered, mx, sharding.whole_job_rate(3, world, 20, mx[0])))
    dist.destroy_process_group()


def test_two_rank_sharding_and_max_reduce():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, mx, rate = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert gathered == [[0, 2, 4, 6], [1, 3, 5]]                      # a partition, round-robin
    assert sorted(sum(gathered, [])) == list(range(7))
    assert mx == [11.0, 3.0]                                          # element-wise max over the two ranks
    assert abs(rate - 2 * 3 * 20 / (11.0 / 1e3)) < 1e-9


def test_single_process_is_identity():
    assert sharding.max_over_ranks([1.5, 2.5]) == [1.5, 2.5]
    assert sharding.shard_images(5, 0, 1) == [0, 1, 2, 3, 4]
    assert sharding.whole_job_rate(3, 1, 10, 20.0) == 1500.0

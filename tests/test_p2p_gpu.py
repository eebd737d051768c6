"""coin_b200.p2p.PeerAllReduce (gradient all-reduce over NVLink peer memory, BASELINE.json configs[3]) against NCCL on two
ranks of one box: whole buffer, repeated calls (epochs), bucket sub-ranges. Needs >= 2 GPUs: skipped on a 1-GPU box."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    from coin_b200 import p2p
    ar = p2p.PeerAllReduce(1_000_003, dev)            # rounded up to a multiple of 4 * world
    assert ar.nelem % (4 * world) == 0 and ar.nelem >= 1_000_003
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    for _ in range(3):
        x = torch.randn(ar.nelem, device=dev, generator=g)
        ar.buffer.copy_(x)
        want = x.clone()
        dist.all_reduce(want)
        ar.all_reduce()
        ar.check()
        assert torch.equal(ar.buffer, want)           # two ranks: a + b in the same order on both sides
    bucket = ar.nelem // 4 // (4 * world) * (4 * world)
    x = torch.randn(ar.nelem, device=dev, generator=g)
    ar.buffer.copy_(x)
    want = x.clone()
    side = torch.cuda.Stream(device=dev, priority=-1)
    side.wait_stream(torch.cuda.current_stream())
    for b in range(4):
        dist.all_reduce(want[b * bucket:(b + 1) * bucket])
        ar.all_reduce(b * bucket, bucket, stream=side)
    torch.cuda.current_stream().wait_stream(side)
    ar.check()
    assert torch.equal(ar.buffer, want)
    with pytest.raises(ValueError):
        ar.all_reduce(1, 8)                           # offset not a multiple of 4 * world
    dist.destroy_process_group()


def test_peer_allreduce_two_ranks_equals_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs of one box")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, 29655), nprocs=2, join=True)

"""torchrun check of coin_b200.p2p.PeerAllReduce (>= 2 GPUs): correctness against the sum computed through NCCL (exact: both
add the ranks' values in order 0..n-1 only when n = 2; otherwise compared to 1e-6 relative), repeated calls, sub-ranges, and
timing alone and beside the ROIAlign grids (configs[3] sizes) next to NCCL's."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coin_b200 import ops, p2p, synth  # noqa: E402
from coin_b200._lib import check, lib  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
nelem = int(200e6 // 4)
ar = p2p.PeerAllReduce(nelem, dev)
g = torch.Generator(device=dev).manual_seed(100 + rank)

# ---- correctness: 3 rounds over the whole buffer, then 8 bucket-sized sub-ranges
ok = True
for it in range(3):
    x = torch.randn(ar.nelem, device=dev, generator=g)
    ar.buffer.copy_(x)
    want = x.clone()
    dist.all_reduce(want)
    ar.all_reduce()
    ar.check()
    err = float((ar.buffer - want).abs().max() / want.abs().max())
    ok &= err < 1e-6
    if rank == 0:
        print(f"round {it}: max rel err vs NCCL {err:.2e}", flush=True)
bucket = ar.nelem // 8 // (4 * world) * (4 * world)
x = torch.randn(ar.nelem, device=dev, generator=g)
ar.buffer.copy_(x)
want = x.clone()
for b in range(8):
    dist.all_reduce(want[b * bucket:(b + 1) * bucket])
    ar.all_reduce(b * bucket, bucket)
ar.check()
err = float((ar.buffer - want).abs().max() / want.abs().max())
ok &= err < 1e-6
if rank == 0:
    print(f"8 buckets: max rel err vs NCCL {err:.2e}", flush=True)

# ---- timing: alone, and beside ROIAlign forward + backward at the BDD shape
shape = synth.SHAPES["bdd_2000"]
gs = synth.gen(rank)
xf = synth.features(gs, shape).to(dev)
n, c, h, w = xf.shape
boxes = [synth.random_boxes(gs, shape.rois, shape.height, shape.width) for _ in range(n)]
rois = torch.cat([torch.cat((torch.full((len(b), 1), float(i)), b), 1) for i, b in enumerate(boxes)]).to(dev)
nhwc = ops.to_nhwc_f32(xf)
go = torch.randn(rois.shape[0], c, 14, 14, device=dev)
buf = torch.zeros((n, h, w, c), device=dev)
lv = ops._levels([buf], (1 / 16,))
nccl_buckets = [torch.randn(int(25e6 // 4), device=dev) for _ in range(8)]
comm = torch.cuda.Stream(device=dev, priority=-1)


def roi():
    ops.roi_align_forward([nhwc], (1 / 16,), rois, None, (14, 14), 0, True, torch.float32)
    check(lib.coin_roi_align_bwd(lv, 1, ops._ptr(rois), ops._ptr(None), ops._ptr(go), 0, c, rois.shape[0], 14, 14, 0, 1, ops._stream()))


def ar_p2p():
    for b in range(8):
        ar.all_reduce(b * bucket, bucket, stream=comm)


def ar_nccl():
    with torch.cuda.stream(comm):
        for b in nccl_buckets:
            dist.all_reduce(b)


def timed(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.current_stream().wait_stream(comm)
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    torch.cuda.current_stream().wait_stream(comm)
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


res = {"roi": timed(roi), "p2p": timed(ar_p2p), "nccl": timed(ar_nccl),
       "roi+p2p": timed(lambda: (ar_p2p(), roi())), "roi+nccl": timed(lambda: (ar_nccl(), roi()))}
ar.check()
if rank == 0:
    nbytes = 8 * bucket * 4
    print({k: round(v, 3) for k, v in res.items()}, f"| p2p bus GB/s {2 * (world - 1) / world * nbytes / res['p2p'] / 1e6:.0f}",
          f"nccl {2 * (world - 1) / world * 200e6 / res['nccl'] / 1e6:.0f}", "| correct" if ok else "| WRONG", flush=True)
dist.destroy_process_group()

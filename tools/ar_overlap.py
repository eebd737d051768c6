"""Can a NCCL all-reduce run BESIDE the ROIAlign grids, or only between them? (torchrun, >= 2 ranks; configs[3] sizes)
Times 10 x ROIAlign forward+backward (BDD shape) alone, 10 x all-reduce of 200 MB alone, and both issued together in the two
possible orders, issued from a NORMAL-priority stream (COIN_BENCH_NCCL_HIPRI only sets the process group's own option).

Measured on 2 x B200 (ms per iteration): roi 2.27, all-reduce 0.56, together 2.81 - 2.82 in either order, with either value of the
process-group option, with NCCL_MAX_NCHANNELS=4 (all-reduce 2.40 alone, 4.12 together) and with the ROIAlign grids capped to 3 or
2 CTAs per SM by shared-memory padding (3.11 -> 3.54, 5.09 -> 5.54): the sum, every time. The ROIAlign grids hold the register
file (4 CTAs x 224 threads x 72 registers = 64.5 k of 65.5 k per SM) and refill every freed slot from their queue of thousands
of CTAs; a collective issued at the same priority waits for the grid to drain. What changes it is the priority of the stream the
collective is ISSUED from: tools/p2p_allreduce_check.py and bench.py --workload bdd_2000 issue it from a high-priority stream and
~70 % of the all-reduce hides under the step (NCCL and the peer-memory kernels of coin_b200/p2p.py alike)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coin_b200 import ops, synth  # noqa: E402
from coin_b200._lib import check, lib  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True) if os.environ.get("COIN_BENCH_NCCL_HIPRI", "1") != "0" else None
dist.init_process_group("nccl", device_id=dev, pg_options=opts)
shape = synth.SHAPES["bdd_2000"]
g = synth.gen(rank)
x = synth.features(g, shape).to(dev)
n, c, h, w = x.shape
boxes = [synth.random_boxes(g, shape.rois, shape.height, shape.width) for _ in range(n)]
rois = torch.cat([torch.cat((torch.full((len(b), 1), float(i)), b), 1) for i, b in enumerate(boxes)]).to(dev)
nhwc = ops.to_nhwc_f32(x)
go = torch.randn(rois.shape[0], c, 14, 14, device=dev)
buf = torch.zeros((n, h, w, c), device=dev)
lv = ops._levels([buf], (1 / 16,))
buckets = [torch.randn(int(25e6 // 4), device=dev) for _ in range(8)]
comm = torch.cuda.Stream(device=dev)


def roi():
    ops.roi_align_forward([nhwc], (1 / 16,), rois, None, (14, 14), 0, True, torch.float32)
    check(lib.coin_roi_align_bwd(lv, 1, ops._ptr(rois), ops._ptr(None), ops._ptr(go), 0, c, rois.shape[0], 14, 14, 0, 1, ops._stream()))


def ar():
    with torch.cuda.stream(comm):
        for b in buckets:
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
    return a.elapsed_time(b) / reps


def both_roi_first():
    roi()
    ar()


def both_ar_first():
    ar()
    roi()


res = {"roi": timed(roi), "allreduce": timed(ar), "roi_then_ar": timed(both_roi_first), "ar_then_roi": timed(both_ar_first)}
if rank == 0:
    print({k: round(v, 3) for k, v in res.items()}, "hipri", opts is not None, flush=True)
dist.destroy_process_group()

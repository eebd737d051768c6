"""Times the default ROIAlign fwd/bwd kernels on the bench shapes under env-var configurations given as
'NAME=VAL,NAME=VAL' arguments (CUDA events; the 1.2 GB output exceeds L2 so no flush is needed)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coin_b200 import _lib, ops, synth  # noqa: E402
from coin_b200._lib import lib, check  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3


def main():
    dev = torch.device("cuda:0")
    cfgs = sys.argv[1:] or [""]
    shapes = (("foggy_roi_head", 14), ("foggy_cpu", 7), ("bdd_2000", 14))
    if os.environ.get("ROI_TIME_SHAPES"):
        shapes = tuple(s for s in shapes if s[0] in os.environ["ROI_TIME_SHAPES"].split(","))
    for name, pooled in shapes:
        shape = synth.SHAPES[name]
        g = synth.gen()
        x = synth.features(g, shape).to(dev)
        n, c, h, w = x.shape
        boxes = [synth.random_boxes(g, shape.rois, shape.height, shape.width) for _ in range(n)]
        rois = torch.cat([torch.cat((torch.full((len(b), 1), float(i)), b), 1) for i, b in enumerate(boxes)]).to(dev)
        k_all = rois.shape[0]
        if os.environ.get("ROI_TIME_SORT"):   # largest RoIs first (or last): how much of the time is the grid's tail?
            area = (rois[:, 3] - rois[:, 1]) * (rois[:, 4] - rois[:, 2])
            mode = os.environ["ROI_TIME_SORT"]
            order = area.argsort(descending=True)
            if mode == "asc":
                order = order.flip(0)
            elif mode == "mix":        # big, small, big, small, ...: ends with the medium ones
                order = torch.stack((order[: k_all // 2], order.flip(0)[: k_all // 2]), 1).flatten()
            elif mode.startswith("tail"):   # input order, but the smallest p % go last (largest of them first)
                p = int(mode[4:] or 20)
                cut = k_all - k_all * p // 100
                small = torch.zeros(k_all, dtype=torch.bool, device=rois.device)
                small[order[cut:]] = True
                idx = torch.arange(k_all, device=rois.device)
                order = torch.cat((idx[~small], order[cut:]))
            rois = rois[order].contiguous()
        k = rois.shape[0]
        nhwc = ops.to_nhwc_f32(x)
        out_bytes = k * c * pooled * pooled * 4
        alg_f = out_bytes + x.numel() * 4 + k * 20
        alg_b = out_bytes + 2 * x.numel() * 4 + k * 20
        io = torch.float16 if os.environ.get("ROI_TIME_FP16") else torch.float32   # fp16 I/O: the autocast path
        esz = 2 if io == torch.float16 else 4
        out_bytes = k * c * pooled * pooled * esz
        alg_f = out_bytes + x.numel() * 4 + k * 20
        alg_b = out_bytes + 2 * x.numel() * 4 + k * 20
        go = torch.randn(k, c, pooled, pooled, device=dev).to(io)
        buf = torch.zeros((n, h, w, c), device=dev)
        lv = ops._levels([buf], (1 / 16,))
        print(f"== {name}: K={k} C={c} map {h}x{w} pooled {pooled}: algorithmic fwd {alg_f/1e6:.1f} MB bwd {alg_b/1e6:.1f} MB", flush=True)
        for cfg in cfgs:
            for kv in filter(None, cfg.split(",")):
                a, b = kv.split("=")
                _lib.set_option(a, int(b))
            tf = timeit(lambda: ops.roi_align_forward([nhwc], (1 / 16,), rois, None, (pooled, pooled), 0, True, io))
            tb = timeit(lambda: check(lib.coin_roi_align_bwd(lv, 1, ops._ptr(rois), ops._ptr(None), ops._ptr(go), 0 if esz == 4 else 1, c, k,
                                                             pooled, pooled, 0, 1, ops._stream())))
            print(f"[{cfg}] fwd {tf:8.1f} us ({alg_f/tf/1e3:7.1f} GB/s)   bwd {tb:8.1f} us ({alg_b/tb/1e3:7.1f} GB/s)", flush=True)
            for kv in filter(None, cfg.split(",")):
                _lib.set_option(kv.split("=")[0], None)


if __name__ == "__main__":
    main()

"""Times ROIAlign fwd/bwd kernel configurations on the bench shape (CUDA events, L2 flushed by size)."""
import itertools
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coin_b200 import ops, synth  # noqa: E402


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
    return s.elapsed_time(e) / iters * 1e3  # us


def main():
    dev = torch.device("cuda:0")
    for name, pooled in (("foggy_roi_head", 14), ("foggy_cpu", 7)):
        shape = synth.SHAPES[name]
        g = synth.gen()
        x = synth.features(g, shape).to(dev)
        n, c, h, w = x.shape
        boxes = [synth.random_boxes(g, shape.rois, shape.height, shape.width) for _ in range(n)]
        rois = torch.cat([torch.cat((torch.full((len(b), 1), float(i)), b), 1) for i, b in enumerate(boxes)]).to(dev)
        k = rois.shape[0]
        nhwc = ops.to_nhwc_f32(x)
        out_bytes = k * c * pooled * pooled * 4
        alg = out_bytes + x.numel() * 4 + k * 20
        print(f"== {name}: K={k} C={c} map {h}x{w} pooled {pooled}: out {out_bytes/1e6:.1f} MB, algorithmic {alg/1e6:.1f} MB")
        t = timeit(lambda: ops.to_nhwc_f32(x))
        print(f"nchw->nhwc: {t:.1f} us")
        go = torch.randn(k, c, pooled, pooled, device=dev)
        for pwc, cpl, rows, exact in itertools.product((7, 14), (2, 4), (1, 2, 4), (0, 1)):
            if pooled == 7 and (pwc == 14 or rows == 4):
                continue
            os.environ["COIN_ROI_FWD_PWC"], os.environ["COIN_ROI_FWD_CPL"] = str(pwc), str(cpl)
            os.environ["COIN_ROI_BWD_PWC"], os.environ["COIN_ROI_BWD_CPL"] = str(pwc), str(cpl)
            os.environ["COIN_ROI_FWD_ROWS"] = os.environ["COIN_ROI_BWD_ROWS"] = str(rows)
            os.environ["COIN_ROI_EXACT"] = str(exact)
            tf = timeit(lambda: ops.roi_align_forward([nhwc], (1 / 16,), rois, None, (pooled, pooled), 0, True, torch.float32))
            buf = torch.zeros((n, h, w, c), device=dev)
            lv = ops._levels([buf], (1 / 16,))
            from coin_b200._lib import lib, check
            def bwd():
                check(lib.coin_roi_align_bwd(lv, 1, ops._ptr(rois), ops._ptr(None), ops._ptr(go), 0, c, k, pooled, pooled, 0, 1, ops._stream()))
            tb = timeit(bwd)
            print(f"pwc={pwc:2d} cpl={cpl} rows={rows} exact={exact}: fwd {tf:8.1f} us ({alg/tf/1e3:7.1f} GB/s)   bwd {tb:8.1f} us ({(out_bytes + 2*x.numel()*4)/tb/1e3:7.1f} GB/s)")
        import torchvision
        tt = timeit(lambda: torchvision.ops.roi_align(x, rois, (pooled, pooled), 1 / 16, 0, True))
        print(f"torchvision CUDA roi_align fwd: {tt:.1f} us ({alg/tt/1e3:.1f} GB/s)")
        xx = x.clone().requires_grad_(True)
        o = torchvision.ops.roi_align(xx, rois, (pooled, pooled), 1 / 16, 0, True)
        tb = timeit(lambda: torch.autograd.grad(o, xx, go, retain_graph=True))
        print(f"torchvision CUDA roi_align bwd: {tb:.1f} us")


if __name__ == "__main__":
    main()

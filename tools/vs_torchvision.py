"""Incumbent-GPU comparator: torchvision 0.26 CUDA roi_align / nms against libcoinops on the bench
shapes, CUDA events, same box. Prints a markdown table."""
import os
import sys

import torch
import torchvision

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import coin_b200  # noqa: E402
from coin_b200 import ops, synth  # noqa: E402
from coin_b200._lib import check, lib  # noqa: E402

dev = torch.device("cuda:0")


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


rows = []
for name in ("foggy_roi_head", "foggy_cpu", "bdd_2000"):
    shape = synth.SHAPES[name]
    p = shape.pooled
    g = synth.gen()
    x = synth.features(g, shape).to(dev)
    n, c, h, w = x.shape
    boxes = [synth.random_boxes(g, shape.rois, shape.height, shape.width) for _ in range(n)]
    rois = torch.cat([torch.cat((torch.full((len(b), 1), float(i)), b), 1) for i, b in enumerate(boxes)]).to(dev)
    k = rois.shape[0]
    nhwc = ops.to_nhwc_f32(x)
    go = torch.randn(k, c, p, p, device=dev)
    buf = torch.zeros((n, h, w, c), device=dev)
    lv = ops._levels([buf], (1 / 16,))
    ours_f = timeit(lambda: ops.roi_align_forward([nhwc], (1 / 16,), rois, None, (p, p), 0, True, torch.float32))
    ours_b = timeit(lambda: check(lib.coin_roi_align_bwd(lv, 1, ops._ptr(rois), ops._ptr(None), ops._ptr(go), 0, c, k, p, p, 0, 1, ops._stream())))
    tv_f = timeit(lambda: torchvision.ops.roi_align(x, rois, (p, p), 1 / 16, 0, True))
    xx = x.clone().requires_grad_(True)
    o = torchvision.ops.roi_align(xx, rois, (p, p), 1 / 16, 0, True)
    tv_b = timeit(lambda: torch.autograd.grad(o, xx, go, retain_graph=True))
    rows.append((f"ROIAlign forward, {name} (K={k}, C={c}, {p}x{p})", tv_f, ours_f))
    rows.append((f"ROIAlign backward, {name}", tv_b, ours_b))
    del o, xx, go, buf
for nbox, thr, mk in ((12000, 0.7, 2000), (6000, 0.7, 1000), (100000, 0.5, -1)):
    g = synth.gen(5)
    b = synth.random_boxes(g, nbox, 600, 1200).to(dev)
    s = torch.randn(nbox, generator=g).to(dev)
    tv = timeit(lambda: torchvision.ops.nms(b, s, thr)[:mk] if mk > 0 else torchvision.ops.nms(b, s, thr), iters=5)
    ours = timeit(lambda: ops.batched_nms(b, s, None, thr, "plain", mk, sync=False), iters=5)
    rows.append((f"NMS {nbox} boxes, thr {thr}, keep[:{mk}] (ours: sync-free, count stays on the device)", tv, ours))
# RPN.predict_proposals for one image and level (decode all anchors, sort, top-k, finite/clip/nonempty, nms, keep[:post]):
# the detectron2 0.5 sequence in eager PyTorch + torchvision CUDA nms against coin_rpn_proposals (one launch chain)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import d2_ref  # noqa: E402  (anchor generation only; the eager sequence below runs on the GPU)
for (hf, wf), pre, post in (((37, 75), 12000, 2000), ((37, 75), 6000, 1000)):
    g = synth.gen(9)
    anchors = d2_ref.grid_anchors(hf, wf, 16, d2_ref.cell_anchors()).to(dev)
    a = anchors.shape[0]
    deltas = (0.2 * torch.randn(a, 4, generator=g)).to(dev)
    logits = torch.randn(a, generator=g).to(dev)
    size = (hf * 16, wf * 16)

    def eager():
        w_, h_ = anchors[:, 2] - anchors[:, 0], anchors[:, 3] - anchors[:, 1]
        cx, cy = anchors[:, 0] + 0.5 * w_, anchors[:, 1] + 0.5 * h_
        dw, dh = deltas[:, 2].clamp(max=4.135166556742356), deltas[:, 3].clamp(max=4.135166556742356)
        pcx, pcy, pw_, ph_ = deltas[:, 0] * w_ + cx, deltas[:, 1] * h_ + cy, torch.exp(dw) * w_, torch.exp(dh) * h_
        props = torch.stack((pcx - 0.5 * pw_, pcy - 0.5 * ph_, pcx + 0.5 * pw_, pcy + 0.5 * ph_), dim=1)
        sl, idx = logits.sort(descending=True)
        sc, bx = sl[:pre], props[idx[:pre]]
        valid = torch.isfinite(bx).all(dim=1) & torch.isfinite(sc)
        if not valid.all():
            bx, sc = bx[valid], sc[valid]
        bx = torch.stack((bx[:, 0].clamp(0, size[1]), bx[:, 1].clamp(0, size[0]), bx[:, 2].clamp(0, size[1]), bx[:, 3].clamp(0, size[0])), 1)
        keep = ((bx[:, 2] - bx[:, 0]) > 0) & ((bx[:, 3] - bx[:, 1]) > 0)
        if keep.sum().item() != len(bx):
            bx, sc = bx[keep], sc[keep]
        k_ = torchvision.ops.nms(bx, sc, 0.7)[:post]
        return bx[k_], sc[k_]

    tv = timeit(eager, iters=5)
    ours = timeit(lambda: ops.rpn_proposals(anchors, deltas, logits, size, pre, post, 0.7, sync=False), iters=5)
    rows.append((f"RPN.predict_proposals, {a} anchors, pre {pre} / post {post} (torch eager + torchvision nms vs one sync-free chain)", tv, ours))
print("| operator | torchvision 0.26 CUDA (us) | libcoinops (us) | speed-up |\n|---|---:|---:|---:|")
for name, tv, ours in rows:
    print(f"| {name} | {tv:.1f} | {ours:.1f} | {tv / ours:.1f}x |")

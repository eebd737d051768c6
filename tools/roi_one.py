"""One ROIAlign fwd+bwd launch set on the bench shape (for ncu)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coin_b200 import ops, synth
from coin_b200._lib import lib, check
dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "foggy_roi_head"
shape = synth.SHAPES[name]; pooled = shape.pooled
g = synth.gen()
x = synth.features(g, shape).to(dev)
n, c, h, w = x.shape
boxes = [synth.random_boxes(g, shape.rois, shape.height, shape.width) for _ in range(n)]
rois = torch.cat([torch.cat((torch.full((len(b), 1), float(i)), b), 1) for i, b in enumerate(boxes)]).to(dev)
k = rois.shape[0]
nhwc = ops.to_nhwc_f32(x)
go = torch.randn(k, c, pooled, pooled, device=dev)
buf = torch.zeros((n, h, w, c), device=dev)
lv = ops._levels([buf], (1 / 16,))
perm = ops.roi_launch_order(rois)      # as in the step: the smallest 20 % of the RoIs are launched last
for _ in range(int(os.environ.get("ITERS", "2"))):
    out = ops.roi_align_forward([nhwc], (1 / 16,), rois, None, (pooled, pooled), 0, True, torch.float32, perm=perm)
    check(lib.coin_roi_align_bwd_ord(lv, 1, ops._ptr(rois), ops._ptr(None), ops._ptr(go), 0, c, k, pooled, pooled, 0, 1, ops._ptr(None), ops._ptr(perm),
                                     ops._stream()))
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))

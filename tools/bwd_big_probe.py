import sys, torch
sys.path.insert(0, '/root/repo')
import coin_b200
from coin_b200 import synth, ops, _lib
dev = torch.device('cuda:0')
g = synth.gen(4711)
x = torch.randn(1, 1024, 37, 75, generator=g).to(dev)
boxes = synth.random_boxes(g, 600, 600, 1200)
big = torch.tensor([[0.0, 0.0, 1200.0, 600.0], [0.0, 431.2, 1200.0, 434.8], [3.0, 0.0, 40.0, 600.0], [0.0, 0.0, 1199.0, 599.0]])
def rois_of(b): return torch.cat((torch.zeros(len(b), 1), b), 1).to(dev)
def timed(fn):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 5 * 1e3
for name, bx in (("plain", boxes), ("mixed", torch.cat((boxes[:300], big, boxes[300:])))):
    r = rois_of(bx)
    go = torch.randn(r.shape[0], 1024, 14, 14, device=dev)
    tb = timed(lambda: ops.roi_align_backward(go, [(1, 1024, 37, 75)], (1 / 16,), r, None, (14, 14), 0, True, [torch.float32]))
    with _lib.options(COIN_ROI_REG=0):
        ts = timed(lambda: ops.roi_align_backward(go, [(1, 1024, 37, 75)], (1 / 16,), r, None, (14, 14), 0, True, [torch.float32]))
    print(name, "bwd reg", round(tb, 1), "us; bwd sep", round(ts, 1), "us")

"""Times the register-tile ROIAlign forward on every private (C) box of one rank's data alone (K = 1 launches) and prints the
slowest with their geometry: which boxes make a CTA run long?  COIN_BENCH_SEED_OFFSET selects the rank."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coin_b200 import _lib, ops, pipeline, synth  # noqa: E402

dev = torch.device("cuda:0")
shape = synth.SHAPES["foggy_roi_head"]
batch = synth.image_batch(shape, seed=synth.SEED + int(os.environ.get("COIN_BENCH_SEED_OFFSET", "0")))
step = pipeline.RoIPathStep(shape, dev)
d = step.to_device(batch)
res = step.finalize(step.run_static(d, backward=False))
nhwc = ops.to_nhwc_f32(d["features"])
rows = []
for i in range(shape.images):
    c = res["abc"][i]["RCNN"][2]["gt_boxes"]
    rows.append(torch.cat((torch.full((len(c), 1), float(i), device=dev), c), 1))
rois = torch.cat(rows)
times = []
with _lib.options(COIN_ROI_REG_MINK=0, COIN_ROI_REG_CHANS_SMALL=256):
    for k in range(rois.shape[0]):
        r = rois[k:k + 1].contiguous()
        for _ in range(2):
            ops.roi_align_forward([nhwc], (1 / 16,), r, None, (14, 14), 0, True, torch.float32)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.roi_align_forward([nhwc], (1 / 16,), r, None, (14, 14), 0, True, torch.float32)
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b) * 1e3)
t = torch.tensor(times)
for k in t.argsort(descending=True)[:8].tolist():
    x = rois[k].tolist()
    print(f"{t[k]:8.1f} us  box {x[1]:.1f} {x[2]:.1f} {x[3]:.1f} {x[4]:.1f}  w {x[3]-x[1]:.1f} h {x[4]-x[2]:.1f}")
print("median", float(t.median()))

"""Step time (ms per graph replay) for the data of ranks 0..7 on ONE GPU: every rank of bench.py draws its own images
(seed + rank) and the step time depends on them through the private-box ROIAlign. Usage: python tools/step_seeds.py [workload]
with the COIN_STEP_* environment switches of coin_b200/pipeline.py."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coin_b200 import pipeline, synth  # noqa: E402

dev = torch.device("cuda:0")
shape = synth.SHAPES[sys.argv[1] if len(sys.argv) > 1 else "foggy_roi_head"]
share = None
for r in range(8):
    batch = synth.image_batch(shape, seed=synth.SEED + r)
    step = pipeline.RoIPathStep(shape, dev, share=share)
    share = share or step
    d = step.to_device(batch)
    out = step.capture(d, backward=True)
    for _ in range(5):
        step.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        step.replay()
    b.record()
    torch.cuda.synchronize()
    res = step.finalize(out)
    print(f"rank-{r} data: {a.elapsed_time(b) / 50:.3f} ms/step   private boxes {res['pooled_c'].shape[0]}", flush=True)
    del step, d, out, res
    torch.cuda.empty_cache()

"""One sync-free step, launched serially (no side streams) after a warm-up: for `ncu --metrics gpu__time_duration.sum`
(per-kernel durations of everything a step launches; tools/ncu_launches.py --sum aggregates)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coin_b200 import pipeline, synth  # noqa: E402

dev = torch.device("cuda:0")
shape = synth.SHAPES[sys.argv[1] if len(sys.argv) > 1 else "foggy_roi_head"]
step = pipeline.RoIPathStep(shape, dev)
step.overlap = False
d = step.to_device(synth.image_batch(shape))
for _ in range(2):
    step.run_static(d, backward=True)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step.run_static(d, backward=True)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()

"""Host-side cost of one pipeline step: cProfile of RoIPathStep.run (steady state) + wall time per step."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coin_b200 import pipeline, synth  # noqa: E402

dev = torch.device("cuda:0")
shape = synth.SHAPES[sys.argv[1] if len(sys.argv) > 1 else "foggy_roi_head"]
batch = synth.image_batch(shape)
step = pipeline.RoIPathStep(shape, dev)
d = step.to_device(batch)
for overlap in (True, False):
    step.overlap = overlap
    for _ in range(5):
        step.run(d)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        step.run(d)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"overlap={overlap}: host issue {1e3*(t1-t0)/20:.2f} ms/step, with final sync {1e3*(t2-t0)/20:.2f} ms/step")
step.overlap = True
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    step.run(d)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)

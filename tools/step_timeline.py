"""Timeline of one graph-replayed step: named external CUDA events captured inside the graph, printed as
offsets (us) from the step's start, averaged over a few replays."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coin_b200 import pipeline, synth  # noqa: E402

dev = torch.device("cuda:0")
shape = synth.SHAPES[sys.argv[1] if len(sys.argv) > 1 else "foggy_roi_head"]
step = pipeline.RoIPathStep(shape, dev)
d = step.to_device(synth.image_batch(shape, seed=synth.SEED + int(os.environ.get("COIN_BENCH_SEED_OFFSET", "0"))))
step.timeline = {}
step.capture(d, backward=True)
tl = step.timeline
acc = {k: 0.0 for k in tl}
n = 10
for _ in range(3):
    step.replay()
for _ in range(n):
    step.replay()
    torch.cuda.synchronize()
    for k, e in tl.items():
        acc[k] += tl["start"].elapsed_time(e) * 1e3
for k, v in sorted(acc.items(), key=lambda kv: kv[1]):
    print(f"{v / n:9.1f} us  {k}")

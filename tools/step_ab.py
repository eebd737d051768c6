"""A/B of library options on the whole graph-replayed step: python tools/step_ab.py "" "NAME=VAL,NAME=VAL" ...
Each configuration is captured afresh (options are read at launch = capture time) and timed over 100 replays (CUDA events),
for the data of ranks 0 and 1; the ROIAlign kernels alone are timed by tools/roi_time.py."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coin_b200 import _lib, pipeline, synth  # noqa: E402

dev = torch.device("cuda:0")
shape = synth.SHAPES[os.environ.get("STEP_AB_SHAPE", "foggy_roi_head")]
share = None
for cfg in sys.argv[1:] or [""]:
    kvs = [kv.split("=") for kv in cfg.split(",") if kv]
    for a, b in kvs:
        _lib.set_option(a, int(b))
    line = []
    for r in range(2):
        batch = synth.image_batch(shape, seed=synth.SEED + r)
        step = pipeline.RoIPathStep(shape, dev, share=share)
        share = share or step
        d = step.to_device(batch)
        out = step.capture(d, backward=True)
        for _ in range(10):
            step.replay()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(100):
            step.replay()
        e.record()
        torch.cuda.synchronize()
        line.append(s.elapsed_time(e) / 100)
        del step, d, out
        torch.cuda.empty_cache()
    for a, _ in kvs:
        _lib.set_option(a, None)
    print(f"[{cfg}] " + "  ".join(f"{t:.4f} ms/step" for t in line), flush=True)

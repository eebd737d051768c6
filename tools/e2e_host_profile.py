"""Host-side time per phase of the double-buffered end-to-end runner (pipeline.PipelinedSteps): is the e2e step bound by the GPU,
the link, or by the Python that packs ~100 result tensors per step?"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coin_b200 import pipeline, synth  # noqa: E402

dev = torch.device("cuda:0")
shape = synth.SHAPES["foggy_roi_head"]
batch = synth.image_batch(shape)
step = pipeline.RoIPathStep(shape, dev)
pinned = step.host_inputs(batch)
d = step.h2d(pinned)
pipe = pipeline.PipelinedSteps(step, d, backward=True)
pipe.load_inputs(pinned)
pipe.run(None, 4)
acc = {"h2d": 0.0, "compute": 0.0, "d2h_wait": 0.0, "d2h_pack": 0.0}
orig_d2h = pipe._d2h


def timed(name, fn):
    def w(*a):
        t = time.perf_counter()
        r = fn(*a)
        acc[name] += time.perf_counter() - t
        return r
    return w


pipe._h2d = timed("h2d", pipe._h2d)
pipe._compute = timed("compute", pipe._compute)


def d2h(n):
    s = n % 2
    t = time.perf_counter()
    pipe.compute_done[s].synchronize()
    acc["d2h_wait"] += time.perf_counter() - t
    t = time.perf_counter()
    orig_d2h(n)
    acc["d2h_pack"] += time.perf_counter() - t


pipe._d2h = d2h
gev = []
for slot in pipe.slots:
    def rep(slot=slot, orig=slot.replay):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = orig()
        b.record()
        gev.append((a, b))
        return r
    slot.replay = rep
steps = 100
torch.cuda.synchronize()
t0 = time.perf_counter()
pipe.run(None, steps)
torch.cuda.synchronize()
total = time.perf_counter() - t0
raw_gaps = [gev[i][1].elapsed_time(gev[i + 1][0]) for i in range(len(gev) - 1)]
print("gaps (ms), steps 40..55:", [round(x, 2) for x in raw_gaps[40:56]], "mean", round(sum(raw_gaps) / len(raw_gaps), 3))
print("graphs (ms), steps 40..55:", [round(a.elapsed_time(b), 2) for a, b in gev[40:56]])
gt = sorted(a.elapsed_time(b) for a, b in gev)
gaps = sorted(raw_gaps)
print(f"e2e {1e3 * total / steps:.3f} ms/step; host per step (ms):", {k: round(1e3 * v / steps, 3) for k, v in acc.items()},
      f"| graph on the GPU: median {gt[len(gt) // 2]:.3f} ms, gap between graphs: median {gaps[len(gaps) // 2]:.3f} ms")

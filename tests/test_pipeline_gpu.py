"""End-to-end parity of one RoI-path step (coin_b200/pipeline.py) against the CPU mirror built from
the oracle, plus the size-independent properties used at BASELINE.json's full sizes."""
import pytest
import torch

from coin_b200 import pipeline, synth
from oracle import pipeline_ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("w_a", [1.0, 0.5])
def test_step_tiny_matches_oracle(dev, w_a):
    shape = synth.SHAPES["tiny"]
    batch = synth.image_batch(shape)
    step = pipeline.RoIPathStep(shape, dev, weight_for_box_a=w_a)
    got = step.run(step.to_device(batch), backward=True)
    want = pipeline_ref.run(batch, backward=True, weight_for_box_a=w_a)
    pipeline_ref.compare(got, want, label_budget=0.0 if w_a == 1.0 else 1e-3)
    assert got["summary"]["dets"] == want["summary"]["dets"]


def test_step_foggy_cpu_config_matches_oracle(dev):
    """BASELINE.json configs[0]: 2 images 600x1200, 8 classes, 512 RoIs/img, 100 cloud dets, 7x7."""
    shape = synth.SHAPES["foggy_cpu"]
    small = synth.Shape(**{**shape.__dict__, "channels": 64})   # full geometry, fewer channels: seconds on CPU
    batch = synth.image_batch(small)
    step = pipeline.RoIPathStep(small, dev)
    got = step.run(step.to_device(batch), backward=True)
    want = pipeline_ref.run(batch, backward=True)
    pipeline_ref.compare(got, want)


def test_step_full_size_properties(dev):
    """configs[1] at full size (3 x 512 RoIs, C=1024, 14x14): properties that need no CPU run."""
    shape = synth.SHAPES["foggy_roi_head"]
    batch = synth.image_batch(shape)
    step = pipeline.RoIPathStep(shape, dev)
    d = step.to_device(batch)
    out = step.run(d, backward=True)
    pooled, grad = out["pooled"], out["grad_features"]
    assert pooled.shape == (shape.images * shape.rois, shape.channels, 14, 14) and bool(torch.isfinite(pooled).all())
    # linearity of ROIAlign in the features: pool(2x) == 2 * pool(x) exactly (power-of-two scaling)
    d2 = dict(d)
    d2["features"] = d["features"] * 2.0
    out2 = step.run(d2, backward=False)
    assert torch.equal(out2["pooled"], pooled * 2.0)
    # adjointness: <pool(x), G> == <x, pool^T(G)> (fp64 accumulation of both inner products)
    lhs = (pooled.double() * step.head_grad.double()).sum()
    rhs = (d["features"].double() * grad.double()).sum()
    assert abs(float(lhs - rhs)) <= 1e-5 * abs(float(lhs))
    # idempotence of NMS: the kept RPN boxes survive a second NMS unchanged
    from coin_b200 import ops
    for i in range(shape.images):
        keep = out["rpn_keep"][i]
        again = ops.nms(d[f"{i}.rpn_boxes"][keep], d[f"{i}.rpn_scores"][keep], 0.7)
        assert torch.equal(again, torch.arange(keep.numel(), device=dev))
        sc = d[f"{i}.rpn_scores"][keep]
        assert bool((sc[:-1] >= sc[1:]).all())
    # every proposal label is consistent with its matched IoU
    for i in range(shape.images):
        idx, lab = out["roi_labels"][i]
        assert int(idx.min()) >= 0 and set(lab.unique().tolist()) <= {-1, 0, 1}


@pytest.mark.parametrize("w_a", [1.0, 0.5])
def test_static_step_and_graph_replay_match_oracle(dev, w_a):
    """The sync-free step (device-side lengths, *_dev entry points): eager, and captured in a CUDA graph and
    replayed on fresh inputs, against the CPU oracle."""
    shape = synth.SHAPES["tiny"]
    batch = synth.image_batch(shape)
    step = pipeline.RoIPathStep(shape, dev, weight_for_box_a=w_a)
    budget = 0.0 if w_a == 1.0 else 1e-3
    want = pipeline_ref.run(batch, backward=True, weight_for_box_a=w_a)
    got = step.finalize(step.run_static(step.to_device(batch), backward=True))
    pipeline_ref.compare(got, want, label_budget=budget)
    assert got["summary"]["dets"] == want["summary"]["dets"]
    # capture on the inputs of ANOTHER batch, then replay on this one: the graph must not bake in any length
    other = synth.image_batch(shape, seed=synth.SEED + 7)
    d_static = step.to_device(other)
    step.capture(d_static, backward=True)
    want_other = pipeline_ref.run(other, backward=True, weight_for_box_a=w_a)
    pipeline_ref.compare(step.finalize(step.replay()), want_other, label_budget=budget)
    step.copy_inputs(step.host_inputs(batch))
    got2 = step.finalize(step.replay())
    pipeline_ref.compare(got2, want, label_budget=budget)
    assert got2["summary"] == got["summary"]


def test_static_step_equals_eager_step_full_size(dev):
    """configs[1] at full size: the graph-replayed sync-free step returns exactly what the eager step returns."""
    shape = synth.SHAPES["foggy_roi_head"]
    batch = synth.image_batch(shape)
    step = pipeline.RoIPathStep(shape, dev)
    d = step.to_device(batch)
    eager = step.run(d, backward=False)
    step.capture(d, backward=False)
    got = step.finalize(step.replay())
    assert got["summary"]["dets"] == eager["summary"]["dets"] and got["summary"]["rpn_keep"] == eager["summary"]["rpn_keep"]
    for i in range(shape.images):
        for k, v in eager["dets"][i].items():
            assert torch.equal(got["dets"][i][k], v), k
        assert torch.equal(got["rpn_keep"][i], eager["rpn_keep"][i])
        for tag in ("RCNN", "RPN"):
            for ge, ee in zip(got["abc"][i][tag], eager["abc"][i][tag]):
                assert (ge is None) == (ee is None)
                if ee is not None:
                    for k, v in ee.items():
                        assert torch.equal(ge[k], v), (i, tag, k)
        for ge, ee in zip(got["roi_labels"][i], eager["roi_labels"][i]):
            assert torch.equal(ge, ee)
        for ge, ee in zip(got["rpn_labels"][i], eager["rpn_labels"][i]):
            assert torch.equal(ge, ee)
    assert torch.equal(got["pooled"], eager["pooled"]) and torch.equal(got["pooled_c"], eager["pooled_c"])


def test_pipelined_end_to_end_steps_match_oracle(dev):
    """Double-buffered end-to-end execution (H2D / graph / D2H of consecutive steps overlapped): both graph
    instances end up holding the oracle's result and the host copies equal the device tensors."""
    shape = synth.SHAPES["tiny"]
    batch = synth.image_batch(shape)
    step = pipeline.RoIPathStep(shape, dev)
    pinned = step.host_inputs(batch)
    pipe = pipeline.PipelinedSteps(step, step.h2d(pinned), backward=True)
    pipe.run(pinned, 5)
    torch.cuda.synchronize()
    want = pipeline_ref.run(batch, backward=True)
    for s, slot in enumerate(pipe.slots):
        got = slot.finalize(slot._graph_out)
        pipeline_ref.compare(got, want)
        for i, t in enumerate(slot.result_tensors(got)):
            assert torch.equal(pipe.host_cache[s][i][: t.numel()], t.reshape(-1).cpu())
    assert pipe.d2h_bytes > 0

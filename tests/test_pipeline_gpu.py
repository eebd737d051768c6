"""End-to-end parity of one RoI-path step (coin_b200/pipeline.py) against the CPU mirror built from
the oracle, plus the size-independent properties used at BASELINE.json's full sizes."""
import pytest
import torch

from coin_b200 import _lib, pipeline, synth
from oracle import pipeline_ref

pytestmark = pytest.mark.gpu


def _exact_on_device_boxes(got, batch):
    """Labelling stage on bit-identical inputs: the oracle's S3/S2 fed with the DEVICE's A/B/C boxes must equal the
    device's labels entry for entry (budget 0)."""
    roi, rpn = pipeline_ref.label_stage(got["abc"], batch)
    for i in range(len(roi)):
        for g, w in zip(got["roi_labels"][i], roi[i]):
            assert torch.equal(g.cpu(), w), f"roi_labels[{i}]"
        for g, w in zip(got["rpn_labels"][i], rpn[i]):
            assert torch.equal(g.cpu(), w), f"rpn_labels[{i}]"


@pytest.mark.parametrize("exact", [False, True], ids=["default_kernels", "bit_exact_roi_align"])
@pytest.mark.parametrize("w_a", [1.0, 0.5])
def test_step_tiny_matches_oracle(dev, w_a, exact):
    """exact=False runs the kernels bench.py times (register-tile / separable ROIAlign); exact=True the parity kernel."""
    shape = synth.SHAPES["tiny"]
    batch = synth.image_batch(shape)
    step = pipeline.RoIPathStep(shape, dev, weight_for_box_a=w_a)
    with _lib.options(COIN_ROI_EXACT=int(exact)):
        got = step.run(step.to_device(batch), backward=True)
    want = pipeline_ref.run(batch, backward=True, weight_for_box_a=w_a)
    flips = {}
    # w_a != 1: the A/B boxes are score-weighted means that involve the CLIP-detector boxes decoded on the device (CUDA
    # expf vs libm: 1 ulp), and the low-quality rule of the anchor Matcher is discontinuous in the coordinates
    pipeline_ref.compare(got, want, label_budget=0.0 if w_a == 1.0 else 1e-3, pooled_exact=exact, flips=flips)
    print("label entries that differ from the CPU run:", {k: v for k, v in flips.items() if v[0]} or "none")
    _exact_on_device_boxes(got, batch)
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


def test_graph_step_full_channels_matches_oracle(dev):
    """BASELINE configs[1] geometry at FULL width (C = 1024, 512 RoIs, 14x14, 12000-box RPN NMS), one image so that the
    CPU mirror takes seconds: the graph-replayed sync-free step - the register-tile ROIAlign forward / backward
    instantiation bench.py times (CS = 1024) - against the oracle, every output."""
    shape = synth.Shape(**{**synth.SHAPES["foggy_roi_head"].__dict__, "images": 1})
    batch = synth.image_batch(shape)
    step = pipeline.RoIPathStep(shape, dev)
    step.capture(step.to_device(batch), backward=True)
    got = step.finalize(step.replay())
    want = pipeline_ref.run(batch, backward=True)
    pipeline_ref.compare(got, want, pooled_exact=False)
    assert got["pooled"].shape == (512, 1024, 14, 14)
    _exact_on_device_boxes(got, batch)


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
    _exact_on_device_boxes(got, batch)
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
    assert torch.equal(got["pooled"], eager["pooled"])
    # the private-box call: device-side RoI count -> separable kernel in the graph, register-tile kernel in the eager step
    torch.testing.assert_close(got["pooled_c"], eager["pooled_c"], rtol=1e-5, atol=1e-5 * float(d["features"].abs().max()))


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
            assert torch.equal(pipe.host_views[s][i], t.reshape(-1).cpu())
    assert pipe.d2h_bytes > 0


def test_distillation_loss_consumers_match_oracle(dev):
    """Row A14: the reference's distillation losses stay in PyTorch; their INPUTS come from this path.
    RoI side (fast_rcnn.py:541-545): KLDiv(log(softmax(scores_c) + 1e-7), gt_probs_c) on the C-box ROIAlign features
    through a fixed head; RPN side (rpn.py:95-98,326-340): teacher_probs = gt_probs_c[:, :-1].sum(1)[matched_idxs]
    on the anchors with distillation label 1. Identical head weights on both sides; <= 1e-5 relative."""
    shape = synth.SHAPES["tiny"]
    batch = synth.image_batch(shape)
    step = pipeline.RoIPathStep(shape, dev)
    got = step.finalize(step.run_static(step.to_device(batch), backward=False))
    want = pipeline_ref.run(batch, backward=False)
    g = synth.gen(91)
    k1 = shape.classes + 1
    w_head = (torch.randn(shape.channels, k1, generator=g) * 0.2).double()
    obj_logits = torch.randn(want["rpn_labels"][0][0].numel(), generator=g).double()
    kl = torch.nn.KLDivLoss(reduction="mean")   # as the reference (fast_rcnn.py:273, rpn.py:15)

    def roi_loss(out, to_cpu):
        feats = to_cpu(out["pooled_c"]).double().mean(dim=(2, 3))                       # [nC, C]
        p = torch.softmax(feats @ w_head, dim=1)
        q = torch.cat([to_cpu(out["abc"][i]["RCNN"][2]["gt_probs"]).double() for i in range(shape.images)])
        return kl(torch.log(p + 1e-7), q)

    def rpn_loss(out, to_cpu, i):
        gt_labels, matched, didx, dlab = (to_cpu(t) for t in out["rpn_labels"][i])
        probs_c = to_cpu(out["abc"][i]["RPN"][2]["gt_probs"]).double()
        if probs_c.shape[0] == 0:
            return torch.zeros((), dtype=torch.float64)
        teacher = probs_c[:, :-1].sum(1)[didx.long()].clamp(0.0, 1.0)   # a sum of fp32 probabilities can exceed 1 by an ulp
        valid = dlab > 0
        p = torch.sigmoid(obj_logits[valid])
        p = torch.stack((p, 1 - p), dim=1)
        q = torch.stack((teacher[valid], 1 - teacher[valid]), dim=1)
        return kl(torch.log(p + 1e-7), q) if bool(valid.any()) else torch.zeros((), dtype=torch.float64)

    cpu = lambda t: t.cpu()
    ident = lambda t: t
    a, b = roi_loss(got, cpu), roi_loss(want, ident)
    assert abs(float(a - b)) <= 1e-5 * abs(float(b)), (float(a), float(b))
    for i in range(shape.images):
        a, b = rpn_loss(got, cpu, i), rpn_loss(want, ident, i)
        assert abs(float(a - b)) <= 1e-5 * max(abs(float(b)), 1e-12), (i, float(a), float(b))

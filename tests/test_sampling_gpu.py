"""SURVEY 8(f) rank 2 on the device: proposal classes + fg/bg subsample (replay of the reference's seeded torch.randperm
draws: bit-exact against the reference's own outputs; device Philox draw: bit-exact against the oracle's restatement of the
same generator), the fused label_and_sample_proposals, RPN._subsample_labels, and the two distillation losses with their
gradients (forward against the reference's outputs / the oracle, backward against torch autograd of the oracle)."""
import pytest
import torch

from coin_b200 import integration, layers, losses, ops, synth
from coin_b200.structures import Boxes, Instances
from conftest import load_golden
from oracle import coin_ref, d2_ref

pytestmark = pytest.mark.gpu
LABELS = load_golden("labels_ref.pt")


def _inst(d, size, dev):
    i = Instances(size)
    for k, v in d.items():
        i.set(k, Boxes(v.to(dev)) if k.endswith("boxes") else v.to(dev))
    return i


def test_label_and_sample_proposals_equals_reference_outputs(dev):
    """clip_roi_heads.py:342-399 ('step_two', PROPOSAL_APPEND_GT, Matcher([0.5],[0,1])) run unmodified with
    torch.manual_seed(2024) -> tests/golden/labels_ref.pt; here the same seed replays the two randperm draws."""
    m = layers.Matcher([0.5], [0, 1], allow_low_quality_matches=False)
    for c in LABELS["roi"]:
        size = (600, 1200)
        p = Instances(size)
        p.proposal_boxes = Boxes(c["proposals"].to(dev))
        n_prop = len(c["proposals"])
        g = synth.gen(0)
        p.objectness_logits = torch.zeros(n_prop, device=dev)
        torch.manual_seed(c["torch_seed"])
        res = integration.label_and_sample_proposals([p], ([_inst(c["a"], size, dev)], [_inst(c["b"], size, dev)],
                                                           [_inst(c["c"], size, dev)]), m, c["num_classes"],
                                                     c["batch_size_per_image"], c["positive_fraction"])
        for got, name in zip(res[0], ("a", "b", "bg")):
            want = c["sampled"][name]
            for k, v in want.items():
                if k == "objectness_logits":
                    continue            # the fixture's proposal logits are random; only their GT-row constant is defined
                gv = got.get(k)
                gv = gv.tensor if isinstance(gv, Boxes) else gv
                assert torch.equal(gv.cpu(), v), (c["label"], name, k)


def test_device_generator_equals_oracle_policy(dev):
    for c in LABELS["roi"]:
        for offset in (0, 5):
            sampled, temp = integration.sample_proposals(c["matched_idxs"].to(dev), c["matched_labels"].to(dev),
                                                         c["gt_classes_cat"].to(dev), c["num_classes"], 512, 0.25,
                                                         generator="device", seed=2024, offset=offset)
            ws, wt = d2_ref.sample_proposals(c["matched_idxs"], c["matched_labels"], c["gt_classes_cat"], c["num_classes"], 512,
                                             0.25, None, 2024, offset)
            assert torch.equal(sampled.cpu(), ws) and torch.equal(temp.cpu(), wt), (c["label"], offset)
    # 41 625 anchor labels, 256 per image, half positive (rpn.py:231): radix-select of 128 keys out of ~40 000
    for r in LABELS["rpn"]:
        lab = r["labels_before_sampling"]
        pos, neg = ops.subsample_labels(lab.to(dev), 256, 0.5, 0, None, seed=99, offset=1)
        wp, wn = d2_ref.subsample_labels(lab, 256, 0.5, 0, None, 99, 1)
        assert torch.equal(pos.cpu(), wp) and torch.equal(neg.cpu(), wn), r["label"]
    # empty and degenerate inputs
    e = torch.empty(0, dtype=torch.int64, device=dev)
    pos, neg = ops.subsample_labels(e, 16, 0.5, 0)
    assert pos.numel() == 0 and neg.numel() == 0
    allneg = torch.zeros(10, dtype=torch.int64, device=dev)
    pos, neg = ops.subsample_labels(allneg, 16, 0.5, 0, seed=1)
    assert pos.numel() == 0 and sorted(neg.tolist()) == list(range(10))
    with pytest.raises(ValueError):
        ops.subsample_labels(allneg, 4, 0.5, 0, perms=(torch.zeros(0, dtype=torch.int64, device=dev),
                                                       torch.full((10,), 11, dtype=torch.int64, device=dev)))


def test_label_and_sample_anchors_equals_reference_outputs(dev):
    """rpn.py:199-254 run unmodified with torch.manual_seed(2024) (labels after RPN._subsample_labels and the
    no-consistent-box epilogue), replayed here with the same seed."""
    m = layers.Matcher([0.3, 0.7], [0, -1, 1], allow_low_quality_matches=True)
    for r in LABELS["rpn"]:
        hf, wf = r["anchors_hw"]
        anchors = Boxes(d2_ref.grid_anchors(hf, wf, 16, d2_ref.cell_anchors()).to(dev))
        torch.manual_seed(r["torch_seed"])
        lab, mgb, didx, dlab = integration.label_and_sample_anchors(m, Boxes(r["a"]["gt_boxes"].to(dev)),
                                                                    Boxes(r["c"]["gt_boxes"].to(dev)), anchors,
                                                                    r["batch_size_per_image"], r["positive_fraction"])
        assert torch.equal(lab.cpu(), r["gt_labels"]), r["label"]
        assert torch.equal(didx.cpu().to(torch.int32), r["all_matched_idxs"]), r["label"]
        assert torch.equal(dlab.cpu(), r["distillation_labels"]), r["label"]
        assert torch.equal(mgb[:256].cpu(), r["matched_gt_boxes_head"]), r["label"]
        assert torch.equal(mgb.double().sum(0).cpu(), r["matched_gt_boxes_colsum"]), r["label"]


PRETRAIN = load_golden("labels_pretrain_ref.pt")


@pytest.mark.parametrize("via_branch", [False, True])
def test_pretrain_label_and_sample_proposals_equals_reference_outputs(dev, via_branch):
    """clip_roi_heads.py:286-340 ('pre_train', PROPOSAL_APPEND_GT, without / with no_thresh_boxes, also without any gt box)
    run unmodified with torch.manual_seed(2024) -> tests/golden/labels_pretrain_ref.pt; every field of the (fg, bg) pair."""
    m = layers.Matcher([0.5], [0, 1], allow_low_quality_matches=False)
    for c in PRETRAIN["roi"]:
        size = tuple(c["image_size"])
        p = Instances(size)
        p.proposal_boxes = Boxes(c["proposals"].to(dev))
        p.objectness_logits = c["objectness_logits"].to(dev)
        t = Instances(size)
        t.gt_boxes = Boxes(c["gt_boxes"].to(dev))
        t.gt_classes_offline = c["gt_classes_offline"].to(dev)
        t.gt_probs = c["gt_probs"].to(dev)
        if c["with_no_thresh"]:
            t._fields["no_thresh_boxes"] = Boxes(c["no_thresh_boxes"].to(dev))    # (its length is its own)
        torch.manual_seed(c["torch_seed"])
        args = ([p], [t], m, c["num_classes"], c["batch_size_per_image"], c["positive_fraction"])
        res = (integration.label_and_sample_proposals(*args, branch="pre_train") if via_branch
               else integration.label_and_sample_proposals_pretrain(*args))
        assert t.has("no_thresh_boxes") == c["with_no_thresh"]
        for got, name in zip(res[0], ("fg", "bg")):
            want = c["sampled"][name]
            assert set(got.get_fields()) == set(want), (c["label"], name)
            for k, v in want.items():
                gv = got.get(k)
                gv = gv.tensor if isinstance(gv, Boxes) else gv
                assert torch.equal(gv.cpu(), v), (c["label"], c["with_no_thresh"], name, k)


def test_pretrain_label_and_sample_anchors_equals_reference_outputs(dev):
    """rpn.py:139-197 ('pre_train') run unmodified with torch.manual_seed(2024), replayed with the same seed."""
    m = layers.Matcher([0.3, 0.7], [0, -1, 1], allow_low_quality_matches=True)
    for r in PRETRAIN["rpn"]:
        hf, wf = r["anchors_hw"]
        anchors = Boxes(d2_ref.grid_anchors(hf, wf, 16, d2_ref.cell_anchors()).to(dev))
        nt = Boxes(r["no_thresh_boxes"].to(dev)) if r["with_no_thresh"] else None
        torch.manual_seed(r["torch_seed"])
        lab, mgb = integration.label_and_sample_anchors_pretrain(m, Boxes(r["gt_boxes"].to(dev)), nt, anchors,
                                                                 r["batch_size_per_image"], r["positive_fraction"])
        what = (r["label"], r["with_no_thresh"])
        assert torch.equal(lab.cpu(), r["gt_labels"]), what
        assert torch.equal(mgb[:256].cpu(), r["matched_gt_boxes_head"]), what
        assert torch.equal(mgb.double().sum(0).cpu(), r["matched_gt_boxes_colsum"]), what


def test_rpn_distillation_loss_and_gradient(dev):
    by_label = {c["label"]: c for c in LABELS["rpn"]}
    for c in LABELS["rpn_loss"]:
        r = by_label[c["label"]]
        hf, wf = r["anchors_hw"]
        logits = torch.randn(1, hf * wf * 15, generator=synth.gen(c["logits_seed"]))[0]
        cc = r["c"]
        idx = r["all_matched_idxs"].long()
        teacher_d = ops.rpn_teacher_probs(cc["gt_probs"].to(dev) if len(cc["gt_boxes"]) else None, idx.to(dev))
        teacher = cc["gt_probs"][:, :-1].sum(1)[idx] if len(cc["gt_boxes"]) else torch.zeros(hf * wf * 15)
        torch.testing.assert_close(teacher_d.cpu(), teacher, rtol=1e-6, atol=1e-7)
        if c["variant"] != "literal":
            teacher, teacher_d = teacher * 0.999, teacher_d * 0.999
        x = logits.to(dev).requires_grad_(True)
        loss = losses.rpn_distillation_loss(x, r["distillation_labels"].to(dev), teacher_d)
        want = c["loss"]
        if "raises" in want:       # 1 - q < 0 by an ulp: NaN here, the reference's assert (rpn.py:343-344) there
            assert bool(torch.isnan(loss))
            continue
        if not want:
            assert float(loss) == 0.0
            loss.backward()
            assert float(x.grad.abs().max()) == 0.0
            continue
        torch.testing.assert_close(loss.cpu(), want["loss_rpn_distillation"], rtol=1e-5, atol=0)
        loss.backward()
        xr = logits.clone().requires_grad_(True)
        coin_ref.rpn_distillation_loss(xr, r["distillation_labels"], teacher).backward()
        torch.testing.assert_close(x.grad.cpu(), xr.grad, rtol=1e-4, atol=1e-5 * float(xr.grad.abs().max()))


@pytest.mark.parametrize("n,live", [(432, None), (600, 137), (1, None), (64, 0)])
def test_roi_distillation_loss_and_gradient(dev, n, live):
    g = synth.gen(900 + n)
    scores = torch.randn(n, 9, generator=g) * 3
    q = torch.softmax(torch.randn(n, 9, generator=g) * 2, dim=1)
    q[::7, -1] = 0.0                                          # exact zeros: xlogy(0, 0) = 0 (cloud probs have no background)
    nd = None if live is None else torch.tensor([live], dtype=torch.int32, device=dev)
    m = n if live is None else live
    x = scores.to(dev).requires_grad_(True)
    loss = losses.roi_distillation_loss(x, q.to(dev), weight=0.5, n_dev=nd)
    if m == 0:
        assert float(loss) == 0.0
        return
    xr = scores[:m].clone().requires_grad_(True)
    want = coin_ref.roi_distillation_loss(xr, q[:m], weight=0.5)
    torch.testing.assert_close(loss.cpu(), want.detach(), rtol=1e-5, atol=0)
    loss.backward()
    want.backward()
    torch.testing.assert_close(x.grad[:m].cpu(), xr.grad, rtol=1e-4, atol=1e-5 * float(xr.grad.abs().max()))
    assert float(x.grad[m:].abs().sum()) == 0.0

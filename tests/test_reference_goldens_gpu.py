"""CUDA path vs outputs of the reference's OWN code (-m gpu).

tests/golden/*_ref.pt hold what the reference's unmodified modules returned on seeded inputs
(tests/golden/make_golden_ref.py; the CPU suite pins the oracle to the same files). Here the device result is
compared with those frozen outputs directly, every call through the C ABI: indices, classes, labels, keep lists,
row order and probabilities bit for bit; box coordinates and fused scores within 1e-5 relative.
"""
import pytest
import torch

import coin_b200
from coin_b200 import integration, layers, ops, synth
from coin_b200.structures import Boxes, Instances
from conftest import load_golden
from oracle import d2_ref

pytestmark = pytest.mark.gpu
PIX_ATOL = 1.2e-4     # 1e-5 relative on ~1e3-px coordinates


def _inst(d, dev, size=(600, 1200)):
    i = Instances(size)
    for k, v in d.items():
        i.set(k, Boxes(v.to(dev)) if k.endswith("boxes") else v.to(dev))
    return i


def _same_fields(got: Instances, want: dict, what, exact_boxes=True):
    assert set(got.get_fields().keys()) == set(want.keys()), (what, sorted(got.get_fields()), sorted(want))
    for k, v in want.items():
        g = got.get(k)
        g = (g.tensor if isinstance(g, Boxes) else g).cpu()
        assert g.shape == v.shape, f"{what}.{k}: shape {tuple(g.shape)} != {tuple(v.shape)}"
        assert g.dtype == v.dtype, f"{what}.{k}: dtype {g.dtype} != {v.dtype}"
        if k.endswith("boxes") and not exact_boxes:
            torch.testing.assert_close(g, v, rtol=1e-5, atol=PIX_ATOL, msg=lambda m: f"{what}.{k}: {m}")
        else:
            assert torch.equal(g, v), f"{what}.{k}: values differ"


ABC = load_golden("abc_ref.pt")


def test_match_dual_teacher_equals_reference_outputs(dev):
    """76 cases of CoinTrainer.match_dual_teacher (trainer.py:338-461) with random.randint pinned to its lower bound:
    A, B, C with every field, in the reference's ROW ORDER (CPython set order replayed on the device). Chain
    clusters, duplicate groups, empty sides, both tags, both box-merging modes. The several-agreeing-duplicates input on
    which the reference fails must fail here too."""
    n_exact = 0
    for c in ABC["cases"]:
        what = f"{c['label']}/{c['tag']}/w_a={c['w_a']}"
        want = c["first"]
        on, off = _inst(c["online"], dev), _inst(c["offline"], dev)
        if isinstance(want, dict):
            with pytest.raises(AssertionError):
                integration.match_dual_teacher(on, off, c["tag"], c["thr"], c["w_a"])
            continue
        got = integration.match_dual_teacher(on, off, c["tag"], c["thr"], c["w_a"])
        for name, g, w in zip("ABC", got, want):
            if w is None:
                assert g is None, what
                continue
            # w_a == 1: boxes are copies of cloud boxes (exact). w_a != 1: a score-weighted mean (two divisions,
            # four multiplies, two adds in fp32, -fmad=false): also bit-identical to the CPU arithmetic
            _same_fields(g, w, f"{what}.{name}", exact_boxes=True)
        n_exact += 1
    assert n_exact == 72


def test_process_and_resize_boxes_equal_reference_outputs(dev):
    g = load_golden("process_ref.pt")
    for c in g["process"]:
        src = _inst(c["in"], dev, c["old_size"])
        keep_in = {k: (v.tensor.clone() if isinstance(v, Boxes) else v.clone()) for k, v in src.get_fields().items()}
        out = integration.process(src, c["old_size"], c["new_size"], c["flip"], c["thresh"], c["keep_name"])
        _same_fields(out, c["out"], f"process/{c['flip']}/{c['thresh']}/{c['keep_name']}")
        for k, v in src.get_fields().items():     # the input is untouched and not aliased (base.py:84 deep-copies)
            t = v.tensor if isinstance(v, Boxes) else v
            assert torch.equal(t, keep_in[k])
            for o in out.get_fields().values():
                o = o.tensor if isinstance(o, Boxes) else o
                assert o.numel() == 0 or o.untyped_storage().data_ptr() != t.untyped_storage().data_ptr(), k
    for c in g["preprocess_results"]:
        results = {"height": 1024, "width": 2048}
        for k, v in c["in"].items():
            results[k] = {"instances": _inst(v, dev, (1024, 2048))}
        out = integration.preprocess_results(results, (600, 1200), "horizontal")
        assert sorted(out.keys()) == c["keys"]
        for tag in ("RCNN", "RPN"):
            _same_fields(out[tag], c["out"][tag], f"preprocess_results/{tag}")
    for c in g["resize_boxes"]:
        got = integration.resize_boxes(c["boxes"].to(dev), c["size"])
        assert torch.equal(got.cpu(), c["out"])
        if len(c["boxes"]):
            assert torch.equal(integration.resize_boxes(c["boxes"].to(dev), c["size"], clip=True).cpu(), c["clipped"])


def test_fast_rcnn_inference_equals_reference_outputs(dev):
    g = load_golden("frcnn_inf_ref.pt")
    for i, c in enumerate(g["cases"]):
        res, kept = integration.fast_rcnn_inference_single_image(c["boxes"].to(dev), c["scores"].to(dev), c["image_shape"],
                                                                 c["score_thresh"], c["nms_thresh"], c["topk"])
        assert torch.equal(kept.cpu(), c["kept"]), i
        _same_fields(res, c["out"], f"frcnn[{i}]")


LABELS = load_golden("labels_ref.pt")


def test_roi_and_anchor_labelling_equal_reference_outputs(dev):
    """The tensors the reference hands to _sample_proposals / _subsample_labels (clip_roi_heads.py:345-362,
    rpn.py:209-228): matched indices, labels with the private-box rule, distillation targets. Budget: zero."""
    m_roi = layers.Matcher([0.5], [0, 1], allow_low_quality_matches=False)
    for c in LABELS["roi"]:
        a, b, cc = (Boxes(c[k]["gt_boxes"].to(dev)) for k in ("a", "b", "c"))
        props = Boxes.cat([Boxes(c["proposals"].to(dev)), a, b])       # add_ground_truth_to_proposals, twice
        idx, lab = integration.label_proposals(m_roi, a, b, cc, props)
        assert torch.equal(idx.cpu(), c["matched_idxs"]) and torch.equal(lab.cpu(), c["matched_labels"]), c["label"]
    m_rpn = layers.Matcher([0.3, 0.7], [0, -1, 1], allow_low_quality_matches=True)
    for c in LABELS["rpn"]:
        hf, wf = c["anchors_hw"]
        anchors = Boxes(d2_ref.grid_anchors(hf, wf, 16, d2_ref.cell_anchors()).to(dev))
        lab, idx0, didx, dlab = integration.label_anchors(m_rpn, Boxes(c["a"]["gt_boxes"].to(dev)),
                                                          Boxes(c["c"]["gt_boxes"].to(dev)), anchors)
        assert torch.equal(lab.cpu(), c["labels_before_sampling"]), c["label"]
        assert torch.equal(didx.cpu().to(torch.int32), c["all_matched_idxs"]), c["label"]
        assert torch.equal(dlab.cpu(), c["distillation_labels"]), c["label"]
        if len(c["a"]["gt_boxes"]):
            mgb = c["a"]["gt_boxes"].to(dev)[idx0]
            assert torch.equal(mgb[:256].cpu(), c["matched_gt_boxes_head"])


def test_rpn_predict_proposals_equals_shimmed_d2(dev):
    for c in LABELS["predict_proposals"]:
        hf, wf = c["hw"]
        g = synth.gen(c["seed"])
        anchors = d2_ref.grid_anchors(hf, wf, 16, d2_ref.cell_anchors())
        deltas = 0.3 * torch.randn(1, anchors.shape[0], 4, generator=g)[0]
        logits = torch.randn(1, anchors.shape[0], generator=g)[0]
        res = integration.rpn_predict_proposals(Boxes(anchors.to(dev)), logits.to(dev), deltas.to(dev), c["image_size"], 0.7,
                                                c["pre"], c["post"])
        assert torch.equal(res.objectness_logits.cpu(), c["objectness_logits"])
        torch.testing.assert_close(res.proposal_boxes.tensor.cpu(), c["proposal_boxes"], rtol=1e-5, atol=PIX_ATOL)


def test_gdino_collection_nms_equals_reference_outputs(dev):
    g = load_golden("gdino_nms_ref.pt")
    for c in g["cases"]:
        nms_module = layers.MyNMS(c["method"])
        out = integration.gdino_collect(_inst(c["in"], dev, (1024, 2048)), nms_module, c["rcnn_thresh"], c["rpn_thresh"],
                                        c["nms_thresh"])
        for tag in ("RCNN", "RPN"):
            want = c["out"][tag]
            got = out[tag]["instances"]
            assert torch.equal(got.pred_classes.cpu(), want["pred_classes"]), (c["method"], tag)
            torch.testing.assert_close(got.pred_boxes.tensor.cpu(), want["pred_boxes"], rtol=1e-5, atol=PIX_ATOL)
            torch.testing.assert_close(got.scores.cpu(), want["scores"], rtol=1e-5, atol=1e-6)
            torch.testing.assert_close(got.probs.cpu(), want["probs"], rtol=1e-5, atol=1e-6)


def test_gdino_collection_with_aug_equals_reference_outputs(dev):
    """gdino_processor.py:287-298 with an 'AUG' set (and an empty ORI set): the three tags."""
    g = load_golden("gdino_aug_ref.pt")
    for c in g["cases"]:
        out = integration.gdino_collect(_inst(c["in"], dev, (1024, 2048)), layers.MyNMS(c["method"]), c["rcnn_thresh"],
                                        c["rpn_thresh"], c["nms_thresh"], aug=_inst(c["aug"], dev, (1024, 2048)))
        for tag in ("RCNN", "RPN", "RPN_AUG"):
            want = c["out"][tag]
            got = out[tag]["instances"]
            assert torch.equal(got.pred_classes.cpu(), want["pred_classes"]), (c["method"], tag)
            torch.testing.assert_close(got.pred_boxes.tensor.cpu(), want["pred_boxes"], rtol=1e-5, atol=PIX_ATOL)
            torch.testing.assert_close(got.scores.cpu(), want["scores"], rtol=1e-5, atol=1e-6)
            torch.testing.assert_close(got.probs.cpu(), want["probs"], rtol=1e-5, atol=1e-6)


def test_unmodified_reference_nms_module_runs_on_this_library(dev):
    """Drop-in proof for coin/layers/nms.py: the fixture outputs came from the reference's MyNMS calling
    detectron2.layers.batched_nms; coin_b200.batched_nms is that symbol's replacement - same keep list on the
    reference's own 'nms' method input."""
    g = load_golden("fusion_nms_nms.pt")
    for c in g["cases"]:
        keep = coin_b200.batched_nms(c["boxes"].to(dev), c["scores"].to(dev), c["labels"].to(dev), c["thr"])
        assert torch.equal(keep.cpu(), c["keep"])

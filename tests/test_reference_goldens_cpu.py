"""Pins the oracle against outputs of the reference's OWN code.

tests/golden/*_ref.pt were produced by tests/golden/make_golden_ref.py, which executes the reference's
unmodified modules (loaded by path from /root/reference behind the detectron2 stand-in of
oracle/d2_shim, see oracle/ref_loader.py). Here every restatement in oracle/coin_ref.py / d2_ref.py that
the GPU parity tests use as the checker is compared with those frozen outputs: all fields, all rows, row
order included, bit for bit (the arithmetic on both sides is torch CPU).
"""
import os
import random

import pytest
import torch

from conftest import load_golden
from oracle import coin_ref, d2_ref
from coin_b200 import synth


def _same_sets(got, want, what):
    for name, gp, wp in zip("ABC", got, want):
        if wp is None:
            assert gp is None, f"{what}: {name} should be None"
            continue
        assert set(gp.keys()) == set(wp.keys()), f"{what}: {name} fields {sorted(gp)} != {sorted(wp)}"
        for k, v in wp.items():
            assert gp[k].shape == v.shape, f"{what}: {name}.{k} shape {tuple(gp[k].shape)} != {tuple(v.shape)}"
            assert gp[k].dtype == v.dtype, f"{what}: {name}.{k} dtype"
            assert torch.equal(gp[k], v), f"{what}: {name}.{k} values"


ABC = load_golden("abc_ref.pt")


def test_abc_golden_inventory():
    """The fixture holds what VERDICT r1 asked for: chain clusters, duplicate groups (matched with / without an
    agreeing member, unmatched, and the several-agreeing-members input on which the reference itself fails), both
    empty-side branches, both tags, both box-merging modes."""
    labels = {c["label"] for c in ABC["cases"]}
    for need in ("chain_and_dups", "chain_shifted_indices", "dup_no_same_class", "dup_two_same_class", "online_empty",
                 "offline_empty", "both_empty", "foggy_cpu.s2024.img0", "tiny.s2024.img1"):
        assert need in labels
    assert {c["tag"] for c in ABC["cases"]} == {"RCNN", "RPN"} and {c["w_a"] for c in ABC["cases"]} == {1.0, 0.5}
    raising = [c for c in ABC["cases"] if isinstance(c["first"], dict)]
    assert raising and all(c["label"] == "dup_two_same_class" for c in raising)


@pytest.mark.parametrize("policy", ["first", "seeded"])
def test_match_dual_teacher_restatement_equals_reference(policy):
    """coin_ref.match_dual_teacher == CoinTrainer.match_dual_teacher (trainer.py:338-461) on 76 cases. 'first':
    random.randint pinned to its lower bound; 'seeded': random.seed(2024) on both sides - equality there shows the
    restatement draws at the same call sites in the same order. Set iteration order is CPython's on both sides."""
    for c in ABC["cases"]:
        what = f"{c['label']}/{c['tag']}/w_a={c['w_a']}/{policy}"
        kw = {"set_order": "cpython"}
        if policy == "seeded":
            random.seed(c["seed"])
            kw["choose"] = coin_ref.random_choice
        want = c[policy]
        if isinstance(want, dict):      # the reference raises on this input (trainer.py:402 / Instances.set)
            with pytest.raises((RuntimeError, AssertionError)):
                coin_ref.match_dual_teacher(c["online"], c["offline"], c["tag"], c["thr"], c["w_a"], **kw)
            continue
        _same_sets(coin_ref.match_dual_teacher(c["online"], c["offline"], c["tag"], c["thr"], c["w_a"], **kw), want, what)


def test_duplicate_and_cluster_helpers_equal_reference():
    """delete_duplicate_boxes (both modes) and filter_result/find_same (util.py:434-482) on every golden input."""
    n_groups = n_clusters = 0
    for h in ABC["helpers"]:
        uniq, groups = coin_ref.delete_duplicate_boxes(h["offline"], return_split=True)
        for k, v in h["uniq"].items():
            assert torch.equal(uniq[k], v), (h["label"], k)
        assert len(groups) == len(h["groups"])
        for g, w in zip(groups, h["groups"]):
            for k, v in w.items():
                assert torch.equal(g[k], v), (h["label"], "group", k)
        merged = coin_ref.delete_duplicate_boxes(h["offline"])
        for k, v in h["merged_first"].items():
            assert torch.equal(merged[k], v), (h["label"], "merged", k)
        n_groups += len(groups)
        if len(h["online"]["gt_classes"]):
            clusters = coin_ref.self_clusters(h["online"]["gt_boxes"], 0.95, set_order="cpython")
            assert len(clusters) == len(h["clusters"]), h["label"]
            for idx, w in zip(clusters, h["clusters"]):
                assert torch.equal(h["online"]["gt_boxes"][torch.tensor(idx)], w["gt_boxes"]), (h["label"], "cluster order")
            n_clusters += len(clusters)
    assert n_groups > 20 and n_clusters > 20


def test_chain_cluster_is_one_component_in_the_reference():
    """A ~ B ~ C ~ D at IoU >= 0.95 with A !~ C: the reference's find_same recursion closes the chain into ONE cluster
    (util.py:459-482), which is what the device's component labelling computes."""
    h = next(x for x in ABC["helpers"] if x["label"] == "chain_and_dups")
    sizes = sorted(len(c["gt_classes"]) for c in h["clusters"])
    assert sizes == [3, 4]
    iou = d2_ref.pairwise_iou(h["online"]["gt_boxes"][:4], h["online"]["gt_boxes"][:4])
    assert float(iou[0, 1]) >= 0.95 and float(iou[0, 2]) < 0.95


def test_process_and_resize_boxes_equal_reference():
    g = load_golden("process_ref.pt")
    for c in g["process"]:
        got = coin_ref.process(c["in"], c["old_size"], c["new_size"], c["flip"], c["thresh"], c["keep_name"])
        assert set(got.keys()) == set(c["out"].keys()), (c["flip"], c["thresh"], c["keep_name"])
        for k, v in c["out"].items():
            assert torch.equal(got[k], v), (c["flip"], c["thresh"], c["keep_name"], k)
        for k, v in c["in_after"].items():           # process() deep-copies: the input is untouched (base.py:84)
            assert torch.equal(c["in"][k], v)
    for c in g["preprocess_results"]:
        got = coin_ref.preprocess_results(c["in"], (1024, 2048), (600, 1200), "horizontal")
        assert c["keys"] == ["RCNN", "RPN", "height", "width"]
        for tag in ("RCNN", "RPN"):
            for k, v in c["out"][tag].items():
                assert torch.equal(got[tag][k], v), (tag, k)
    for c in g["resize_boxes"]:
        got = coin_ref.resize_boxes(c["boxes"], c["size"])
        assert torch.equal(got, c["out"])
        if len(got):
            assert torch.equal(d2_ref.box_clip(got, c["size"]), c["clipped"])


def test_fast_rcnn_inference_equals_reference():
    g = load_golden("frcnn_inf_ref.pt")
    sizes = []
    for c in g["cases"]:
        res, kept = coin_ref.fast_rcnn_inference_single_image(c["boxes"], c["scores"], c["image_shape"], c["score_thresh"],
                                                              c["nms_thresh"], c["topk"])
        assert torch.equal(kept, c["kept"])
        for k, v in c["out"].items():
            assert torch.equal(res[k], v), k
        sizes.append(len(kept))
    assert 0 in sizes and max(sizes) >= 100      # the all-NaN input keeps nothing; topk = -1 keeps everything


LABELS = load_golden("labels_ref.pt")


def test_roi_labelling_equals_reference():
    """add_ground_truth_to_proposals + pairwise_iou + Matcher([0.5],[0,1]) + the C-box relabel, as executed by
    OpenVocabularyRes5ROIHeads.label_and_sample_proposals (clip_roi_heads.py:345-362): the tensors the reference hands
    to _sample_proposals."""
    for c in LABELS["roi"]:
        a, b, cc = c["a"], c["b"], c["c"]
        gt = torch.cat((a["gt_boxes"], b["gt_boxes"], cc["gt_boxes"]))
        props = torch.cat((c["proposals"], a["gt_boxes"], b["gt_boxes"]))
        idx, lab = d2_ref.Matcher([0.5], [0, 1], False)(d2_ref.pairwise_iou(gt, props))
        lab = coin_ref.relabel_roi(idx, lab, len(a["gt_boxes"]), len(b["gt_boxes"]), len(cc["gt_boxes"]))
        assert torch.equal(idx, c["matched_idxs"]), c["label"]
        assert torch.equal(lab, c["matched_labels"]), c["label"]
        assert torch.equal(torch.cat((a["gt_classes"], b["gt_classes_online"], cc["gt_classes"])), c["gt_classes_cat"])


def test_anchor_labelling_equals_reference():
    """pairwise_iou + Matcher([0.3,0.7],[0,-1,1], low quality) + the C-box relabel / distillation targets, as executed
    by DualTeacherRPN.label_and_sample_anchors (rpn.py:209-228): labels before _subsample_labels, all_matched_idxs,
    distillation_labels."""
    for c in LABELS["rpn"]:
        hf, wf = c["anchors_hw"]
        anchors = d2_ref.grid_anchors(hf, wf, 16, d2_ref.cell_anchors())
        gt = torch.cat((c["a"]["gt_boxes"], c["c"]["gt_boxes"]))
        idx, lab = d2_ref.Matcher([0.3, 0.7], [0, -1, 1], True)(d2_ref.pairwise_iou(gt, anchors))
        lab, idx0, dist_idx, dist_lab = coin_ref.relabel_rpn(idx, lab, len(c["a"]["gt_boxes"]), len(c["c"]["gt_boxes"]))
        assert torch.equal(lab, c["labels_before_sampling"]), c["label"]
        assert torch.equal(dist_idx.to(torch.int32), c["all_matched_idxs"]), c["label"]
        assert torch.equal(dist_lab, c["distillation_labels"]), c["label"]
        if len(c["a"]["gt_boxes"]):
            mgb = c["a"]["gt_boxes"][idx0]
            assert torch.equal(mgb[:256], c["matched_gt_boxes_head"])
            assert torch.equal(mgb.double().sum(0), c["matched_gt_boxes_colsum"])


PRETRAIN = load_golden("labels_pretrain_ref.pt")


def test_pretrain_roi_labelling_equals_reference():
    """The 'pre_train' branch of label_and_sample_proposals (clip_roi_heads.py:286-317), without and with
    `no_thresh_boxes`: the (matched_idxs, matched_labels) the reference hands to _sample_proposals, and the sampled
    (fg, bg) sets replayed with the same torch seed."""
    for c in PRETRAIN["roi"]:
        what = (c["label"], c["with_no_thresh"])
        gt, nt = c["gt_boxes"], c["no_thresh_boxes"]
        props = torch.cat((c["proposals"], gt))                    # PROPOSAL_APPEND_GT: only the gt rows join
        rows = torch.cat((gt, nt)) if c["with_no_thresh"] else gt
        idx, lab = d2_ref.Matcher([0.5], [0, 1], False)(d2_ref.pairwise_iou(rows, props))
        if c["with_no_thresh"]:
            lab, idx = coin_ref.relabel_pretrain(idx, lab, len(gt), len(nt))
        assert torch.equal(idx, c["matched_idxs"]), what
        assert torch.equal(lab, c["matched_labels"]), what
        torch.manual_seed(c["torch_seed"])
        k = c["num_classes"]
        pre = c["gt_classes_offline"][idx].clone() if len(gt) else torch.zeros_like(idx) + k
        if len(gt):
            pre[lab == 0], pre[lab == -1] = k, -1
        perms = (torch.randperm(int(((pre != -1) & (pre != k)).sum())), torch.randperm(int((pre == k).sum())))
        sampled, cls = d2_ref.sample_proposals(idx, lab, c["gt_classes_offline"], k, c["batch_size_per_image"],
                                               c["positive_fraction"], perms)
        bg = cls == c["num_classes"]
        assert torch.equal(props[sampled[~bg]], c["sampled"]["fg"]["proposal_boxes"]), what
        assert torch.equal(props[sampled[bg]], c["sampled"]["bg"]["proposal_boxes"]), what
        assert torch.equal(cls[bg], c["sampled"]["bg"]["gt_classes"]), what
        m = idx[sampled][~bg]
        assert torch.equal(gt[m], c["sampled"]["fg"]["gt_boxes"]), what
        assert torch.equal(c["gt_classes_offline"][m], c["sampled"]["fg"]["gt_classes_offline"]), what
        assert torch.equal(c["gt_probs"][m], c["sampled"]["fg"]["gt_probs"]), what


def test_pretrain_anchor_labelling_equals_reference():
    """The 'pre_train' branch of label_and_sample_anchors (rpn.py:139-197): labels before _subsample_labels, labels after
    it and the no-gt epilogue (rpn.py:183-190), matched gt boxes."""
    for c in PRETRAIN["rpn"]:
        what = (c["label"], c["with_no_thresh"])
        hf, wf = c["anchors_hw"]
        anchors = d2_ref.grid_anchors(hf, wf, 16, d2_ref.cell_anchors())
        gt, nt = c["gt_boxes"], c["no_thresh_boxes"]
        rows = torch.cat((gt, nt)) if c["with_no_thresh"] else gt
        idx, lab = d2_ref.Matcher([0.3, 0.7], [0, -1, 1], True)(d2_ref.pairwise_iou(rows, anchors))
        bg_nt = None
        if c["with_no_thresh"]:
            bg_nt = (idx >= len(gt)) & (idx < len(gt) + len(nt)) & (lab == 0)
            lab, idx = coin_ref.relabel_pretrain(idx, lab, len(gt), len(nt))
        assert torch.equal(lab, c["labels_before_sampling"]), what
        torch.manual_seed(c["torch_seed"])
        perms = (torch.randperm(int((lab == 1).sum())), torch.randperm(int((lab == 0).sum())))
        pos, neg = d2_ref.subsample_labels(lab, c["batch_size_per_image"], c["positive_fraction"], 0, perms)
        out = torch.full_like(lab, -1)
        out[pos], out[neg] = 1, 0
        if len(gt) == 0:
            if bg_nt is None:
                out[:] = -1
            else:
                out[~bg_nt] = -1
            mgb = torch.zeros_like(anchors)
        else:
            mgb = gt[idx]
        assert torch.equal(out, c["gt_labels"]), what
        assert torch.equal(mgb[:256], c["matched_gt_boxes_head"]), what
        assert torch.equal(mgb.double().sum(0), c["matched_gt_boxes_colsum"]), what


def test_rpn_distillation_loss_equals_reference():
    by_label = {c["label"]: c for c in LABELS["rpn"]}
    finite = 0
    for c in LABELS["rpn_loss"]:
        r = by_label[c["label"]]
        hf, wf = r["anchors_hw"]
        logits = torch.randn(1, hf * wf * 15, generator=synth.gen(c["logits_seed"]))[0]
        cc = r["c"]
        if len(cc["gt_boxes"]):
            teacher = cc["gt_probs"][:, :-1].sum(1)[r["all_matched_idxs"].long()]
        else:
            teacher = torch.zeros(hf * wf * 15)
        if c["variant"] != "literal":
            teacher = teacher * 0.999
        want = c["loss"]
        if "raises" in want:            # 1 - q < 0 by an ulp: the reference's own NaN assert fires (rpn.py:343-344)
            with pytest.raises(AssertionError):
                coin_ref.rpn_distillation_loss(logits, r["distillation_labels"], teacher)
            continue
        got = coin_ref.rpn_distillation_loss(logits, r["distillation_labels"], teacher)
        if not want:
            assert got is None
        else:
            assert torch.equal(got, want["loss_rpn_distillation"]), c["label"]
            finite += 1
    assert finite >= 2


def test_predict_proposals_equals_shimmed_d2():
    """d2_ref.predict_proposals_single == detectron2's RPN.predict_proposals as restated in oracle/d2_shim (two
    independent restatements of the 0.5 source: the per-image function used by the GPU tests and the batched,
    multi-level form the reference calls at rpn.py:113)."""
    for c in LABELS["predict_proposals"]:
        hf, wf = c["hw"]
        g = synth.gen(c["seed"])
        anchors = d2_ref.grid_anchors(hf, wf, 16, d2_ref.cell_anchors())
        deltas = 0.3 * torch.randn(1, anchors.shape[0], 4, generator=g)[0]
        logits = torch.randn(1, anchors.shape[0], generator=g)[0]
        boxes, sc = d2_ref.predict_proposals_single(anchors, deltas, logits, c["image_size"], 0.7, c["pre"], c["post"])
        assert torch.equal(boxes, c["proposal_boxes"]) and torch.equal(sc, c["objectness_logits"])


def test_gdino_collect_equals_reference():
    g = load_golden("gdino_nms_ref.pt")
    for c in g["cases"]:
        got = coin_ref.gdino_collect(c["in"], c["method"], c["rcnn_thresh"], c["rpn_thresh"], c["nms_thresh"])
        for tag in ("RCNN", "RPN"):
            for k, v in c["out"][tag].items():
                assert torch.equal(got[tag][k], v), (c["method"], tag, k)


def test_gdino_collect_with_aug_equals_reference():
    """post_process with an 'AUG' set (gdino_processor.py:295-297): RPN_AUG = nms(cat(nms(RPN), AUG)); also an empty ORI set."""
    g = load_golden("gdino_aug_ref.pt")
    for c in g["cases"]:
        got = coin_ref.gdino_collect(c["in"], c["method"], c["rcnn_thresh"], c["rpn_thresh"], c["nms_thresh"], aug=c["aug"])
        for tag in ("RCNN", "RPN", "RPN_AUG"):
            for k, v in c["out"][tag].items():
                assert torch.equal(got[tag][k], v), (c["method"], tag, k)


@pytest.mark.skipif(not os.path.isdir("/root/reference/coin"), reason="the reference tree exists in the build container only")
def test_reference_modules_load_unmodified_under_the_stub_finder():
    """oracle/ref_loader.py: every module is the reference's file (by path), not a copy."""
    from oracle import ref_loader
    for name, rel in ref_loader.REAL_COIN.items():
        mod = ref_loader.load(name)
        assert os.path.samefile(mod.__file__, os.path.join("/root/reference", rel))

"""CPU mirror of coin_b200/pipeline.py built from the oracle functions (TEST INFRASTRUCTURE ONLY).

``run`` is also what ``bench.py`` times as the reference arm / cpu_baseline: the reference's own
CPU path for this step is torch + torchvision CPU operators behind detectron2 plus the Python loops
of coin/engine/trainer.py:338-485 -- which is exactly what the functions called here restate.
"""
import math

import torch
import torchvision

from . import coin_ref, d2_ref

ORIG_SCALE = 2048.0 / 1200.0
BBOX_WEIGHTS = (10.0, 10.0, 5.0, 5.0)


def anchors_for(shape):
    hf, wf = shape.feat_hw
    return d2_ref.grid_anchors(hf, wf, shape.stride, d2_ref.cell_anchors())


def head_grad(shape, seed=2024):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed + 1)
    return torch.randn(shape.images * shape.rois, shape.channels, shape.pooled, shape.pooled, generator=g)


def label_stage(abc_per_image, batch, anchors=None):
    """S3 + S2 alone (clip_roi_heads.py:345-362, rpn.py:209-228) on GIVEN pseudo-label sets: fed with the device's own
    A/B/C boxes it shows the labelling exact on bit-identical inputs, with no tolerance budget."""
    sh = batch["shape"]
    anchors = anchors_for(sh) if anchors is None else anchors
    roi_labels, rpn_labels = [], []
    for img, per_tag in zip(batch["images"], abc_per_image):
        a, b, c = (x["gt_boxes"].cpu() for x in per_tag["RCNN"])
        gt = torch.cat((a, b, c))
        props = torch.cat((img["proposals"], a, b))
        idx, lab = d2_ref.Matcher([0.5], [0, 1], False)(d2_ref.pairwise_iou(gt, props))
        roi_labels.append((idx, coin_ref.relabel_roi(idx, lab, len(a), len(b), len(c))))
        a2, c2 = per_tag["RPN"][0]["gt_boxes"].cpu(), per_tag["RPN"][2]["gt_boxes"].cpu()
        idx2, lab2 = d2_ref.Matcher([0.3, 0.7], [0, -1, 1], True)(d2_ref.pairwise_iou(torch.cat((a2, c2)), anchors))
        rpn_labels.append(coin_ref.relabel_rpn(idx2, lab2, len(a2), len(c2)))
    return roi_labels, rpn_labels


def run(batch, backward=True, weight_for_box_a=1.0, anchors=None, grad=None, roi_limit=None):
    """roi_limit: bound the number of RoIs per image fed to ROIAlign (bench.py's bounded CPU sample)."""
    sh = batch["shape"]
    img_size = (sh.height, sh.width)
    feats = batch["features"]
    anchors = anchors_for(sh) if anchors is None else anchors
    t = d2_ref.Box2BoxTransform(BBOX_WEIGHTS)
    out = {"dets": [], "abc": [], "roi_labels": [], "rpn_labels": [], "rpn_keep": []}
    c_rois, rois = [], []
    for i, img in enumerate(batch["images"]):
        dec = t.apply_deltas(img["teacher_deltas"], img["teacher_rois"])
        det, kept = coin_ref.fast_rcnn_inference_single_image(dec, img["teacher_probs"], img_size, 0.05, 0.5, 100)
        det["roi_index"] = kept
        out["dets"].append(det)
        cloud = dict(img["cloud"])
        cloud["gt_boxes"] = coin_ref.process_boxes(img["cloud"]["gt_boxes"] * ORIG_SCALE,
                                                   (sh.height * ORIG_SCALE, sh.width * ORIG_SCALE), img_size, "no")
        clip = {"gt_boxes": det["pred_boxes"], "gt_classes": det["pred_classes"], "scores": det["scores"],
                "probs": det["probs"]}
        per_tag = {tag: coin_ref.match_dual_teacher(cloud, clip, tag, 0.5, weight_for_box_a) for tag in ("RCNN", "RPN")}
        out["abc"].append(per_tag)
        out["rpn_keep"].append(d2_ref.nms(img["rpn_boxes"], img["rpn_scores"], 0.7)[: sh.rpn_post_nms])

        a, b, c = per_tag["RCNN"]
        gt = torch.cat((a["gt_boxes"], b["gt_boxes"], c["gt_boxes"]))
        props = torch.cat((img["proposals"], a["gt_boxes"], b["gt_boxes"]))
        idx, lab = d2_ref.Matcher([0.5], [0, 1], False)(d2_ref.pairwise_iou(gt, props))
        out["roi_labels"].append((idx, coin_ref.relabel_roi(idx, lab, len(a["gt_boxes"]), len(b["gt_boxes"]),
                                                            len(c["gt_boxes"]))))
        a2, _, c2 = per_tag["RPN"]
        gt2 = torch.cat((a2["gt_boxes"], c2["gt_boxes"]))
        idx2, lab2 = d2_ref.Matcher([0.3, 0.7], [0, -1, 1], True)(d2_ref.pairwise_iou(gt2, anchors))
        out["rpn_labels"].append(coin_ref.relabel_rpn(idx2, lab2, len(a2["gt_boxes"]), len(c2["gt_boxes"])))
        cb = c["gt_boxes"]
        c_rois.append(torch.cat((torch.full((cb.shape[0], 1), float(i)), cb), dim=1))
        r = img["rois"] if roi_limit is None else img["rois"][:roi_limit]
        rois.append(torch.cat((torch.full((r.shape[0], 1), float(i)), r), dim=1))

    rois = torch.cat(rois)
    size = (sh.pooled, sh.pooled)
    x = feats.clone().requires_grad_(backward)
    out["pooled"] = torchvision.ops.roi_align(x, rois, size, 1.0 / sh.stride, 0, True)
    with torch.no_grad():
        out["pooled_c"] = torchvision.ops.roi_align(feats, torch.cat(c_rois), size, 1.0 / sh.stride, 0, True)
    if backward:
        g = head_grad(sh) if grad is None else grad
        if roi_limit is not None:
            g = g.view(sh.images, sh.rois, *g.shape[1:])[:, :roi_limit].reshape(-1, *g.shape[1:])
        out["pooled"].backward(g)
        out["grad_features"] = x.grad
        out["pooled"] = out["pooled"].detach()
    out["summary"] = {"dets": [len(d["scores"]) for d in out["dets"]],
                      "rpn_keep": [len(k) for k in out["rpn_keep"]]}
    return out


def _eq(a, b, what):
    a = a.cpu() if hasattr(a, "cpu") else a
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} != {tuple(b.shape)}"
    assert torch.equal(a, b), f"{what}: integer / exact mismatch"


def _close(a, b, what, atol):
    a = a.cpu()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} != {tuple(b.shape)}"
    torch.testing.assert_close(a, b, rtol=1e-5, atol=atol, msg=lambda m: f"{what}: {m}")


def _labels(a, b, what, budget, flips=None):
    a = a.cpu()
    assert a.shape == b.shape, f"{what}: shape"
    bad = int((a != b).sum())
    if flips is not None:
        flips[what] = (bad, a.numel())
    assert bad <= budget * a.numel(), f"{what}: {bad} of {a.numel()} entries differ (budget {budget})"


def compare(got, want, pix_atol=1.2e-4, label_budget=0.0, pooled_exact=None, flips=None):
    """Bit-exact on every index / label / keep list and on ROIAlign forward; 1e-5 relative on floats
    (pixel coordinates: atol 1.2e-4 px; gradients: atol 1e-5 * max|grad|).

    flips: optional dict that receives, per label vector, (number of differing entries, length).
    label_budget: fraction of proposal / anchor labels allowed to differ. It must be 0 whenever the
    pseudo boxes fed to the Matcher are bit-identical on both sides (WEIGHT_FOR_BOX_A == 1: they are the
    cloud boxes). With score-weighted merging the A/B boxes inherit the one-ulp difference between the
    CPU and CUDA exp() of the decode, and the low-quality rule (label 1 for every anchor whose IoU EQUALS
    the row maximum) is discontinuous in the box coordinates."""
    for i, (g, w) in enumerate(zip(got["dets"], want["dets"])):
        _eq(g["roi_index"], w["roi_index"], f"dets[{i}].roi_index")
        _eq(g["pred_classes"], w["pred_classes"], f"dets[{i}].pred_classes")
        _close(g["pred_boxes"], w["pred_boxes"], f"dets[{i}].pred_boxes", pix_atol)
        _eq(g["scores"], w["scores"], f"dets[{i}].scores")
        _eq(g["probs"], w["probs"], f"dets[{i}].probs")
    for i, (g, w) in enumerate(zip(got["rpn_keep"], want["rpn_keep"])):
        _eq(g, w, f"rpn_keep[{i}]")
    for i, (g, w) in enumerate(zip(got["abc"], want["abc"])):
        for tag in ("RCNN", "RPN"):
            for name, gp, wp in zip("ABC", g[tag], w[tag]):
                if wp is None:
                    assert gp is None
                    continue
                for k, v in wp.items():
                    if k == "gt_boxes":
                        _close(gp[k], v, f"abc[{i}].{tag}.{name}.{k}", pix_atol)
                    else:
                        _eq(gp[k], v, f"abc[{i}].{tag}.{name}.{k}")
    for i, (g, w) in enumerate(zip(got["roi_labels"], want["roi_labels"])):
        _labels(g[0], w[0], f"roi_labels[{i}].matched_idxs", label_budget, flips)
        _labels(g[1], w[1], f"roi_labels[{i}].matched_labels", label_budget, flips)
    for i, (g, w) in enumerate(zip(got["rpn_labels"], want["rpn_labels"])):
        for j, name in enumerate(("gt_labels", "matched_idxs", "distillation_idxs", "distillation_labels")):
            _labels(g[j], w[j], f"rpn_labels[{i}].{name}", label_budget, flips)
    if pooled_exact is None:
        pooled_exact = False
    if pooled_exact:   # parity mode of the forward kernel: bit-identical to the CPU kernel
        _eq(got["pooled"], want["pooled"], "pooled (ROIAlign forward, bit-exact)")
    else:              # default FMA accumulation: 1e-5 relative to the feature scale
        _close(got["pooled"], want["pooled"], "pooled", 1e-5 * float(want["pooled"].abs().max()))
    # the C boxes include CLIP-detector boxes whose decoded coordinates differ by an ulp between the CPU
    # and CUDA exp(): the sample positions move by ~1e-5 cell, the pooled value by ~1e-5 * |feature|
    _close(got["pooled_c"], want["pooled_c"], "pooled_c", 1e-5 * float(want["pooled_c"].abs().max()))
    if "grad_features" in want:
        _close(got["grad_features"], want["grad_features"], "grad_features",
               1e-5 * float(want["grad_features"].abs().max()))

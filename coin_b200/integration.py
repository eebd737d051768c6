"""Host-side mirror of the COIN functions that sit on the RoI path, running on device tensors.

Same names / argument meaning as the reference so the parity tests read like calls into it:
    process                          coin/engine/base.py:80-126      (box rescale + flip, field renames)
    fast_rcnn_inference_single_image coin/modeling/roi_heads/fast_rcnn.py:116-175
    match_dual_teacher / merge_boxes coin/engine/trainer.py:338-461,480-485
    delete_duplicate_boxes           coin/utils/util.py:434-457 (device de-dup inside match_abc)
    label_proposals / label_anchors  the pairwise_iou + Matcher + relabel blocks of
                                     coin/modeling/roi_heads/clip_roi_heads.py:351-362 and
                                     coin/modeling/proposal_generator/rpn.py:209-228
The reference moves the teacher detections to the CPU for the matching (trainer.py:469) and back
(:457-459); here everything stays on the device and each function is one or a few launches.
"""
import math
from typing import Optional, Tuple

import torch

from . import ops
from .layers import Matcher
from .structures import Boxes, Instances


def process(instances: Instances, old_size, new_size, random_flip: str, thresh: Optional[float] = None,
            keep_name: bool = False) -> Instances:
    """BASE_Trainer.process: scale boxes from the original image frame into the network-input frame,
    optionally flip, rename pred_* -> gt_* (one kernel for the box math)."""
    img_h, img_w = old_size
    net_h, net_w = new_size
    if random_flip not in ("no", "horizontal", "vertical"):
        raise NotImplementedError
    out = Instances((net_h, net_w))
    fields = dict(instances.get_fields())
    name = "pred_boxes" if "pred_boxes" in fields else "gt_boxes"
    boxes = Boxes(ops.boxes_scale_flip(fields.pop(name).tensor, net_w / img_w, net_h / img_h, random_flip,
                                       (net_h, net_w)))
    for k, v in fields.items():
        # the reference deep-copies every field (base.py:84): the result must not alias its input (e.g. a cache entry)
        out.set(k, v.clone() if hasattr(v, "clone") else v)
    out.set(name if (keep_name or name == "gt_boxes") else "gt_boxes", boxes)
    if not keep_name:
        out.set("gt_classes", out.get("pred_classes"))
        out.remove("pred_classes")
    if thresh is not None:
        return out[instances.scores >= thresh]
    return out


def preprocess_results(results: dict, new_image_size, random_flip: str, thresh: Optional[float] = None) -> dict:
    """BASE_Trainer.preprocess_results (base.py:128-136): process() for both tags of a cached cloud result; a collected
    'RPN_AUG' set replaces 'RPN'. Like the reference it rewrites the dict it is given."""
    old = (results["height"], results["width"])
    results["RCNN"] = process(results["RCNN"]["instances"], old, new_image_size, random_flip, thresh)
    if "RPN_AUG" in results:
        del results["RPN"]
        results["RPN"] = process(results["RPN_AUG"]["instances"], old, new_image_size, random_flip, thresh)
        del results["RPN_AUG"]
    else:
        results["RPN"] = process(results["RPN"]["instances"], old, new_image_size, random_flip, thresh)
    return results


def resize_boxes(boxes: torch.Tensor, size: Tuple[int, int], clip: bool = False) -> torch.Tensor:
    """GDINO.resize_boxes (gdino.py:144-160): cxcywh in [0,1] -> xyxy in pixels of an (H, W) image; clip=True also
    applies the Boxes.clip(size) the caller does next (gdino.py:135-136). One launch instead of a Python loop per box."""
    if boxes.shape[0] == 0:
        return boxes
    return ops.boxes_cxcywh_to_xyxy(boxes, size, clip)


def gdino_collect(ori: Instances, nms_module, rcnn_thresh: float, rpn_thresh: float, nms_thresh: float,
                  aug: Optional[Instances] = None) -> dict:
    """GDINO_PROCESSOR.post_process without ZOOM (gdino_processor.py:287-298) + GDINO_PROCESSOR.nms (:164-182):
    the two score thresholds, then ``mynms.nms`` (coin_b200.layers.MyNMS) per tag. ``aug`` = the detections of the augmented
    view (cfg.INPUT.TEACHER_CLOUD.COLLECT_AUG): adds 'RPN_AUG' = the NMS of the NMS'ed RPN set followed by them (:295-297),
    which ``preprocess_results`` then uses in place of 'RPN'."""
    def nms(boxes, scores, probs, classes):
        _, boxes, scores, probs, labels = nms_module.nms(boxes, scores, probs, classes, nms_thresh)
        inst = Instances(ori.image_size)
        inst.pred_boxes = Boxes(boxes)
        inst.scores = scores
        inst.pred_classes = labels
        inst.probs = probs
        return {"instances": inst}

    out = {}
    for tag, thr in (("RCNN", rcnn_thresh), ("RPN", rpn_thresh)):
        sub = ori[ori.scores >= thr]
        out[tag] = nms(sub.pred_boxes.tensor, sub.scores, sub.probs, sub.pred_classes)
    if aug is not None:
        rpn = out["RPN"]["instances"]
        out["RPN_AUG"] = nms(torch.cat((rpn.pred_boxes.tensor, aug.pred_boxes.tensor)), torch.cat((rpn.scores, aug.scores)),
                             torch.cat((rpn.probs, aug.probs)), torch.cat((rpn.pred_classes, aug.pred_classes)))
    return out


def fast_rcnn_inference_single_image(boxes, scores, image_shape: Tuple[int, int], score_thresh: float,
                                     nms_thresh: float, topk_per_image: int):
    """Returns (Instances(pred_boxes, scores, probs, pred_classes), kept RoI indices)."""
    b, s, p, c, roi = ops.det_postprocess(boxes, scores, image_shape, score_thresh, nms_thresh, topk_per_image)
    result = Instances(image_shape)
    result.pred_boxes = Boxes(b)
    result.scores = s
    result.probs = p
    result.pred_classes = c
    return result, roi


def _gather(inst: Instances, idx: torch.Tensor, names) -> dict:
    return {n: inst.get(n)[idx] for n in names}


def match_dual_teacher(online: Instances, offline: Instances, tag: str, iou_threshold: float = 0.5,
                       weight_for_box_a: float = 1.0, device=None):
    """CoinTrainer.match_dual_teacher for one image and one tag. ``online`` = cloud detections after
    process() (fields gt_boxes, gt_classes, scores, probs), ``offline`` = CLIP-detector detections
    (same fields). Returns (A, B or None, C) Instances with the reference's field names.

    Policy for the reference's non-deterministic choices (random.randint, set iteration order):
    first element / ascending index; see DESIGN.md."""
    nc, nd = len(online), len(offline)
    r = ops.match_abc(online.gt_boxes.tensor, online.gt_classes, online.scores,
                      offline.gt_boxes.tensor, offline.gt_classes, offline.scores, tag, iou_threshold,
                      weight_for_box_a)
    # in the empty-side branches both members of a pair index the same (non-empty) set
    on_src = offline if nc == 0 else online
    off_src = online if nd == 0 else offline
    size = online.image_size

    def pack(on_idx, off_idx, boxes, split_classes):
        out = Instances(size)
        out.gt_boxes = Boxes(boxes)
        if split_classes:
            out.gt_classes_offline = off_src.gt_classes[off_idx]
            out.gt_classes_online = on_src.gt_classes[on_idx]
        else:
            out.gt_classes = off_src.gt_classes[off_idx]
        out.gt_scores_online = on_src.scores[on_idx]
        out.gt_scores_offline = off_src.scores[off_idx]
        out.gt_probs_online = on_src.probs[on_idx]
        out.gt_probs_offline = off_src.probs[off_idx]
        return out

    a = pack(r["a_on"], r["a_off"], r["a_boxes"], False)
    b = pack(r["b_on"], r["b_off"], r["b_boxes"], True) if tag == "RCNN" else None

    # C rows: CLIP-detector rows first (c_off), then cloud rows (c_on)
    off_rows, on_rows = r["c_off"], r["c_on"]
    c = Instances(size)
    fields = ("gt_boxes", "gt_classes", "scores", "probs")
    parts = []
    if nd > 0:
        parts.append({"gt_boxes": offline.gt_boxes.tensor[off_rows], "gt_classes": offline.gt_classes[off_rows],
                      "scores": offline.scores[off_rows], "probs": offline.probs[off_rows]})
    if nc > 0:
        parts.append({"gt_boxes": online.gt_boxes.tensor[on_rows], "gt_classes": online.gt_classes[on_rows],
                      "scores": online.scores[on_rows], "probs": online.probs[on_rows]})
    if not parts:
        parts.append({"gt_boxes": online.gt_boxes.tensor, "gt_classes": online.gt_classes, "scores": online.scores,
                      "probs": online.probs})
    cat = {k: torch.cat([p[k] for p in parts], dim=0) for k in fields}
    c.gt_boxes = Boxes(cat["gt_boxes"])
    c.gt_classes = cat["gt_classes"]
    c.gt_scores = cat["scores"]
    c.gt_probs = cat["probs"]
    if device is not None:
        a, c = a.to(device), c.to(device)
        b = b.to(device) if b is not None else None
    return a, b, c


def label_proposals(matcher: Matcher, a_boxes: Boxes, b_boxes: Boxes, c_boxes: Boxes, proposals: Boxes):
    """clip_roi_heads.py:351-362: IoU of proposals against A|B|C pseudo boxes, Matcher, and the rule
    that foreground matches on a private (C) box are ignored. Returns (matched_idxs, matched_labels)."""
    gt = Boxes.cat([a_boxes, b_boxes, c_boxes])
    idx, lab = matcher.match_boxes(gt, proposals)
    la, lb, lc = len(a_boxes), len(b_boxes), len(c_boxes)
    ops.relabel_roi_(idx, lab, la + lb, la + lb + lc)
    return idx, lab


def sample_proposals(matched_idxs: torch.Tensor, matched_labels: torch.Tensor, gt_classes: torch.Tensor, num_classes: int,
                     batch_size_per_image: int, positive_fraction: float, generator="torch", seed: int = 0, offset: int = 0):
    """detectron2 ``ROIHeads._sample_proposals`` (<- clip_roi_heads.py:317,363): proposal classes (background / ignore rules)
    and the fg / bg subsample, both on the device. Returns (sampled_idxs, gt_classes[sampled_idxs]).

    generator="torch": the reference's draw - two ``torch.randperm`` calls on the global generator, over the positive and
    the negative count (one count read-back), replayed on the device: bit-exact against a seeded reference run.
    generator="device": Philox4x32-10 keyed by (seed, offset) on the device, no host round trip."""
    cls = ops.proposal_classes(matched_idxs, matched_labels, gt_classes, num_classes)
    perms = None
    if generator == "torch":
        n_pos, n_neg = ops.subsample_labels(cls, batch_size_per_image, positive_fraction, num_classes, count_only=True)
        perms = (torch.randperm(n_pos).to(cls.device), torch.randperm(n_neg).to(cls.device))   # same order as subsample_labels
    elif generator != "device":
        raise ValueError(generator)
    fg, bg = ops.subsample_labels(cls, batch_size_per_image, positive_fraction, num_classes, perms, seed, offset)
    sampled = torch.cat([fg, bg], dim=0)
    return sampled, cls[sampled]


def subsample_anchor_labels(label: torch.Tensor, batch_size_per_image: int, positive_fraction: float, generator="torch",
                            seed: int = 0, offset: int = 0) -> torch.Tensor:
    """detectron2 ``RPN._subsample_labels`` (<- rpn.py:231): keeps a random subset of the positive (1) and negative (0)
    anchor labels and sets every other anchor to ignore (-1), in place like the original."""
    perms = None
    if generator == "torch":
        n_pos, n_neg = ops.subsample_labels(label, batch_size_per_image, positive_fraction, 0, count_only=True)
        perms = (torch.randperm(n_pos).to(label.device), torch.randperm(n_neg).to(label.device))
    pos_idx, neg_idx = ops.subsample_labels(label, batch_size_per_image, positive_fraction, 0, perms, seed, offset)
    label.fill_(-1)
    label.scatter_(0, pos_idx, 1)
    label.scatter_(0, neg_idx, 0)
    return label


def label_and_sample_proposals(proposals: list, targets, matcher: Matcher, num_classes: int, batch_size_per_image: int,
                               positive_fraction: float, proposal_append_gt: bool = True, bg_train: bool = True,
                               generator="torch", seed: int = 0, offset: int = 0, branch: str = "step_two") -> list:
    """``OpenVocabularyRes5ROIHeads.label_and_sample_proposals`` (clip_roi_heads.py:283-399). 'step_one' / 'step_two'
    (:342-399): ``proposals`` = one Instances(proposal_boxes, ...) per image; ``targets`` = (A, B, C) lists
    of pseudo-label Instances as ``match_dual_teacher`` returns them. Returns one (proposals_a, proposals_b, proposals_bg)
    triple per image with the reference's fields: the fused IoU + Matcher + private-box rule (A6-A8), the class assignment
    and the subsample run in kernels; the per-group field gathers are index selections. 'pre_train' (:286-340): see
    ``label_and_sample_proposals_pretrain``."""
    if branch == "pre_train":
        return label_and_sample_proposals_pretrain(proposals, targets, matcher, num_classes, batch_size_per_image,
                                                   positive_fraction, proposal_append_gt, generator, seed, offset)
    if branch not in ("step_one", "step_two"):
        raise ValueError(branch)
    a_targets, b_targets, c_targets = targets
    out = []
    for i, (p, a, b, c) in enumerate(zip(proposals, a_targets, b_targets, c_targets)):
        boxes = p.proposal_boxes
        logits = p.objectness_logits if p.has("objectness_logits") else None
        if proposal_append_gt:       # add_ground_truth_to_proposals(a), then (b): their boxes join the proposal list
            boxes = Boxes.cat([boxes, a.gt_boxes, b.gt_boxes])
            if logits is not None:   # ... with the logit of probability 1 - 1e-10 (d2 proposal_utils.py)
                gt_logit = math.log((1.0 - 1e-10) / (1 - (1.0 - 1e-10)))
                logits = torch.cat([logits, torch.full((len(a) + len(b),), gt_logit, dtype=logits.dtype, device=logits.device)])
        len_a, len_b = len(a), len(b)
        idx, lab = label_proposals(matcher, a.gt_boxes, b.gt_boxes, c.gt_boxes, boxes)
        gt_cat = torch.cat([a.gt_classes, b.gt_classes_online, c.gt_classes])
        sampled, temp = sample_proposals(idx, lab, gt_cat, num_classes, batch_size_per_image, positive_fraction, generator,
                                         seed, offset + i)
        m = idx[sampled]
        bg = temp == num_classes
        mask_a = (m >= 0) & (m < len_a) & ~bg
        mask_b = (m >= len_a) & (m < len_a + len_b) & ~bg
        size = p.image_size
        pa, pb, pbg = Instances(size), Instances(size), Instances(size)
        pa.proposal_boxes, pb.proposal_boxes = Boxes(boxes.tensor[sampled[mask_a]]), Boxes(boxes.tensor[sampled[mask_b]])
        pbg.proposal_boxes = Boxes(boxes.tensor[sampled[bg]])
        if logits is not None:
            pa.objectness_logits, pb.objectness_logits = logits[sampled[mask_a]], logits[sampled[mask_b]]
            pbg.objectness_logits = logits[sampled[bg]]
        pbg.gt_classes = temp[bg]
        if not bg_train:
            pbg = pbg[0:0]
        for name, val in a.get_fields().items():
            if name.startswith("gt_"):
                pa.set(name, val[m[mask_a]])
        for name, val in b.get_fields().items():
            if name.startswith("gt_"):
                pb.set(name, val[m[mask_b] - len_a])
        out.append((pa, pb, pbg))
    return out


def label_and_sample_proposals_pretrain(proposals: list, targets: list, matcher: Matcher, num_classes: int,
                                        batch_size_per_image: int, positive_fraction: float, proposal_append_gt: bool = True,
                                        generator="torch", seed: int = 0, offset: int = 0) -> list:
    """The 'pre_train' branch of ``label_and_sample_proposals`` (clip_roi_heads.py:286-340): ``targets`` = one Instances per
    image with ``gt_boxes``, ``gt_classes_offline`` (+ any other gt_* field) and, optionally on ALL of them, ``no_thresh_boxes``
    - boxes that are matched like ground truth but must be neither foreground nor sampled as such (:300-309: label -1 unless
    the Matcher said background, index reset to 0). Returns one (proposals_fg, proposals_bg) pair per image. Unlike the
    reference (:289-291) the targets are not modified: ``no_thresh_boxes`` stays on them."""
    with_nt = len(targets) > 0 and all(t.has("no_thresh_boxes") for t in targets)
    out = []
    for i, (p, t) in enumerate(zip(proposals, targets)):
        boxes = p.proposal_boxes
        logits = p.objectness_logits if p.has("objectness_logits") else None
        gt = t.gt_boxes
        if proposal_append_gt:
            boxes = Boxes.cat([boxes, gt])
            if logits is not None:
                gt_logit = math.log((1.0 - 1e-10) / (1 - (1.0 - 1e-10)))
                logits = torch.cat([logits, torch.full((len(gt),), gt_logit, dtype=logits.dtype, device=logits.device)])
        if with_nt:
            nt = t.get("no_thresh_boxes")
            idx, lab = matcher.match_boxes(Boxes.cat([gt, nt]), boxes)
            lab, idx, _, _ = ops.relabel_rpn_(idx, lab, len(gt), len(nt))      # the same rule as the RPN's C boxes
        else:
            idx, lab = matcher.match_boxes(gt, boxes)
        sampled, temp = sample_proposals(idx, lab, t.gt_classes_offline, num_classes, batch_size_per_image, positive_fraction,
                                         generator, seed, offset + i)
        m = idx[sampled]
        bg = temp == num_classes
        fg = ~bg
        size = p.image_size
        pf, pbg = Instances(size), Instances(size)
        pf.proposal_boxes, pbg.proposal_boxes = Boxes(boxes.tensor[sampled[fg]]), Boxes(boxes.tensor[sampled[bg]])
        if logits is not None:
            pf.objectness_logits, pbg.objectness_logits = logits[sampled[fg]], logits[sampled[bg]]
        pbg.gt_classes = temp[bg]
        for name, val in t.get_fields().items():
            if name.startswith("gt_"):
                pf.set(name, val[m[fg]])
        out.append((pf, pbg))
    return out


def label_anchors(matcher: Matcher, a_boxes: Boxes, c_boxes: Boxes, anchors: Boxes):
    """rpn.py:209-228: returns (gt_labels, matched_idxs, distillation_idxs, distillation_labels)."""
    gt = Boxes.cat([a_boxes, c_boxes])
    idx, lab = matcher.match_boxes(gt, anchors)
    return ops.relabel_rpn_(idx, lab, len(a_boxes), len(c_boxes))


def label_and_sample_anchors(matcher: Matcher, a_boxes: Boxes, c_boxes: Boxes, anchors: Boxes, batch_size_per_image: int,
                             positive_fraction: float, generator="torch", seed: int = 0, offset: int = 0):
    """``DualTeacherRPN.label_and_sample_anchors`` for one image of the 'step_one' / 'step_two' branches (rpn.py:199-254,
    anchor_boundary_thresh < 0): returns (gt_labels, matched_gt_boxes, all_matched_idxs, distillation_labels)."""
    lab, idx, didx, dlab = label_anchors(matcher, a_boxes, c_boxes, anchors)
    before = lab.clone() if len(a_boxes) == 0 else None
    lab = subsample_anchor_labels(lab, batch_size_per_image, positive_fraction, generator, seed, offset)
    if len(a_boxes) == 0:
        # rpn.py:244-248: without consistent boxes only the anchors that matched a private box as background keep a label
        matched_gt_boxes = torch.zeros_like(anchors.tensor)
        if len(c_boxes) == 0:
            lab.fill_(-1)
        else:
            lab[before != 0] = -1
    else:
        matched_gt_boxes = a_boxes.tensor[idx]
    return lab, matched_gt_boxes, didx, dlab


def label_and_sample_anchors_pretrain(matcher: Matcher, gt_boxes: Boxes, no_thresh_boxes: Optional[Boxes], anchors: Boxes,
                                      batch_size_per_image: int, positive_fraction: float, generator="torch", seed: int = 0,
                                      offset: int = 0):
    """``DualTeacherRPN.label_and_sample_anchors`` for one image of the 'pre_train' branch (rpn.py:139-197,
    anchor_boundary_thresh < 0): returns (gt_labels, matched_gt_boxes). ``no_thresh_boxes`` (may be None) are treated exactly
    like the private (C) boxes of the later branches - label -1 unless background, index reset to 0, and without any gt box
    only the anchors that matched one of them as background keep a label (:183-190)."""
    nt = no_thresh_boxes if no_thresh_boxes is not None else Boxes(gt_boxes.tensor.new_zeros((0, 4)))
    lab, mgb, _, _ = label_and_sample_anchors(matcher, gt_boxes, nt, anchors, batch_size_per_image, positive_fraction,
                                              generator, seed, offset)
    return lab, mgb


class AnchorGrid(tuple):
    """One level of detectron2's ``DefaultAnchorGenerator`` (<- rpn.py:64) in closed form: (cell_anchors [ncell,4], Hf, Wf,
    stride, offset). ``rpn_predict_proposals`` accepts it in place of the materialised ``Boxes`` of all anchors."""

    def __new__(cls, cell_anchors, hf: int, wf: int, stride: int, offset: float = 0.0):
        return super().__new__(cls, (torch.as_tensor(cell_anchors, dtype=torch.float32).reshape(-1, 4), int(hf), int(wf),
                                     float(stride), float(offset)))

    @staticmethod
    def cell_anchors(sizes=(32, 64, 128, 256, 512), ratios=(0.5, 1.0, 2.0)) -> torch.Tensor:
        """DefaultAnchorGenerator.generate_cell_anchors (Base-Cloud.yaml:22-23: 5 sizes x 3 ratios)."""
        out = []
        for s in sizes:
            area = s ** 2.0
            for r in ratios:
                w = math.sqrt(area / r)
                h = r * w
                out.append([-w / 2.0, -h / 2.0, w / 2.0, h / 2.0])
        return torch.tensor(out, dtype=torch.float32)

    def materialise(self) -> torch.Tensor:
        """The [Hf*Wf*ncell, 4] array grid_anchors builds (for callers that still want it, e.g. anchor labelling)."""
        cell, hf, wf, stride, offset = self
        sx = torch.arange(offset * stride, wf * stride, step=stride, dtype=torch.float32)
        sy = torch.arange(offset * stride, hf * stride, step=stride, dtype=torch.float32)
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        shifts = torch.stack((xx.reshape(-1), yy.reshape(-1), xx.reshape(-1), yy.reshape(-1)), dim=1)
        return (shifts.view(-1, 1, 4) + cell.view(1, -1, 4)).reshape(-1, 4)


def rpn_predict_proposals(anchors, pred_objectness_logits: torch.Tensor, pred_anchor_deltas: torch.Tensor,
                          image_size: Tuple[int, int], nms_thresh: float, pre_nms_topk: int, post_nms_topk: int,
                          min_box_size: float = 0.0, training: bool = False,
                          weights=(1.0, 1.0, 1.0, 1.0)) -> Instances:
    """detectron2 0.5 ``RPN.predict_proposals`` for one image and one feature level, as reached from
    coin/modeling/proposal_generator/rpn.py:64,113: ``_decode_proposals`` + ``find_top_rpn_proposals``. Returns
    ``Instances(image_size)`` with ``proposal_boxes`` and ``objectness_logits`` like the reference. One launch chain on
    the device (coin_rpn_proposals) and a single 8-byte read-back of (count, status)."""
    if isinstance(anchors, AnchorGrid):     # DefaultAnchorGenerator's grid, generated inside the decode kernel
        boxes, logits, status = ops.rpn_proposals(None, pred_anchor_deltas, pred_objectness_logits, image_size, pre_nms_topk,
                                                  post_nms_topk, nms_thresh, min_box_size, weights, grid=tuple(anchors))
    else:
        boxes, logits, status = ops.rpn_proposals(anchors.tensor if isinstance(anchors, Boxes) else anchors, pred_anchor_deltas,
                                                  pred_objectness_logits, image_size, pre_nms_topk, post_nms_topk, nms_thresh,
                                                  min_box_size, weights)
    if status & 1 and training:
        raise FloatingPointError("Predicted boxes or scores contain Inf/NaN. Training has diverged.")
    res = Instances(image_size)
    res.proposal_boxes = Boxes(boxes)
    res.objectness_logits = logits
    return res


def detector_postprocess(results: Instances, output_height: int, output_width: int) -> Instances:
    """detectron2 0.5 ``modeling/postprocessing.py::detector_postprocess`` as reached from
    ``GeneralizedRCNN._postprocess`` (coin/modeling/meta_arch/clip_rcnn.py:424, clip_rcnn_oracle.py:251): rescale the
    boxes from the network input size to the requested output size, clip, drop empty boxes (the
    scale and clip run in the A13 / A4 kernels). Like detectron2 it rescales the Boxes object the input shares."""
    new_size = (output_height, output_width)
    scale_x, scale_y = output_width / results.image_size[1], output_height / results.image_size[0]
    results = Instances(new_size, **results.get_fields())
    if results.has("pred_boxes"):
        output_boxes = results.pred_boxes
    elif results.has("proposal_boxes"):
        output_boxes = results.proposal_boxes
    else:
        output_boxes = None
    assert output_boxes is not None, "Predictions must contain boxes!"
    output_boxes.scale(scale_x, scale_y)
    output_boxes.clip(results.image_size)
    return results[output_boxes.nonempty()]


def box_reg_loss(box2box_transform, proposal_boxes: torch.Tensor, gt_boxes: torch.Tensor, pred_deltas: torch.Tensor,
                 gt_classes: torch.Tensor, num_classes: int, smooth_l1_beta: float = 0.0, normalizer=None) -> torch.Tensor:
    """``FastRCNNOutputLayers.box_reg_loss`` (coin/modeling/roi_heads/fast_rcnn.py:601-646, smooth_l1 type): the
    regression targets of the foreground proposals come from the A5 kernel (``Box2BoxTransform.get_deltas``, no
    gradient), the smooth-L1 / L1 sum and its normalisation by the number of regions stay differentiable PyTorch."""
    box_dim = proposal_boxes.shape[1]
    fg_inds = torch.nonzero((gt_classes >= 0) & (gt_classes < num_classes), as_tuple=True)[0]
    if pred_deltas.shape[1] == box_dim:   # class-agnostic regression (Base-Cloud.yaml:39)
        fg_pred_deltas = pred_deltas[fg_inds]
    else:
        fg_pred_deltas = pred_deltas.view(-1, num_classes, box_dim)[fg_inds, gt_classes[fg_inds]]
    if fg_inds.numel():
        with torch.no_grad():
            target = box2box_transform.get_deltas(proposal_boxes[fg_inds].contiguous(), gt_boxes[fg_inds].contiguous())
    else:
        target = fg_pred_deltas.detach()
    n = torch.abs(fg_pred_deltas - target)
    if smooth_l1_beta < 1e-5:             # fvcore smooth_l1_loss: plain L1 below this beta
        loss = n.sum()
    else:
        loss = torch.where(n < smooth_l1_beta, 0.5 * n ** 2 / smooth_l1_beta, n - 0.5 * smooth_l1_beta).sum()
    if normalizer is not None:
        return loss / normalizer
    return loss / max(gt_classes.numel(), 1.0)

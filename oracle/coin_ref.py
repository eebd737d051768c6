"""CPU restatement of the hot-path arithmetic the COIN reference authors itself.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py). All tensors are CPU tensors.

A detection set is a plain ``dict`` of equally long CPU tensors (the reference uses detectron2
``Instances``, which cannot be imported here). Field names are the reference's.

Two places of the reference are not functions of their inputs alone:
  * ``random.randint`` picks (coin/engine/trainer.py:385,387; coin/utils/util.py:450);
  * iteration order of CPython ``set`` objects (trainer.py:369,391; util.py:471-482).
Both are exposed as policies: ``choose`` (callable n -> index; default = first = the reference with
random.randint pinned to its lower bound, which is what the device does; pass ``random_choice`` for
the seeded behaviour) and ``set_order`` ("cpython", the default = literal ``list(set)`` as the
reference executes it, which the device replays exactly; "ascending" = sorted, kept to show where the
two differ).
"""
import random
from typing import Callable, Dict, List, Optional, Tuple

import torch

from . import d2_ref

DetSet = Dict[str, torch.Tensor]


# ------------------------------------------------------------------------------------------------
# detection-set helpers (stand-ins for detectron2 Instances indexing / cat)
# ------------------------------------------------------------------------------------------------
def take(d: DetSet, idx) -> DetSet:
    if isinstance(idx, int):
        idx = torch.tensor([idx], dtype=torch.int64)
    return {k: v[idx] for k, v in d.items()}


def cat(sets: List[DetSet]) -> DetSet:
    return {k: torch.cat([s[k] for s in sets], dim=0) for k in sets[0].keys()}


def length(d: DetSet) -> int:
    return int(next(iter(d.values())).shape[0])


def first_choice(n: int) -> int:
    return 0


def random_choice(n: int) -> int:
    return random.randint(0, n - 1)


def _ordered(s: set, set_order: str) -> List[int]:
    return sorted(s) if set_order == "ascending" else list(s)


# ------------------------------------------------------------------------------------------------
# A13  coin/engine/base.py:80-126  box rescale + flip into the network-input frame
# ------------------------------------------------------------------------------------------------
def process_boxes(boxes: torch.Tensor, old_size, new_size, flip: str = "no") -> torch.Tensor:
    img_h, img_w = old_size
    net_h, net_w = new_size
    b = d2_ref.box_scale(boxes.float(), net_w / img_w, net_h / img_h)
    if flip == "horizontal":
        f = b.clone()
        f[:, 0] = net_w - b[:, 2]
        f[:, 2] = net_w - b[:, 0]
        b = f
    elif flip == "vertical":
        f = b.clone()
        f[:, 1] = net_h - b[:, 3]
        f[:, 3] = net_h - b[:, 1]
        b = f
    elif flip != "no":
        raise NotImplementedError(flip)
    return b


def process(inst: DetSet, old_size, new_size, flip: str = "no", thresh=None, keep_name: bool = False) -> DetSet:
    """BASE_Trainer.process, base.py:80-126, on a field dict: deep copy, Boxes.scale, flip, pred_* -> gt_* renames
    (unless keep_name), optional `scores >= thresh` row filter (on the ORIGINAL scores, base.py:117)."""
    out = {k: v.clone() for k, v in inst.items()}
    key = "pred_boxes" if "pred_boxes" in out else "gt_boxes"
    boxes = process_boxes(out.pop(key), old_size, new_size, flip)
    out[key if (keep_name or key == "gt_boxes") else "gt_boxes"] = boxes
    if not keep_name:
        out["gt_classes"] = out.pop("pred_classes")
    if thresh is not None:
        keep = inst["scores"] >= thresh
        out = {k: v[keep] for k, v in out.items()}
    return out


def preprocess_results(results: Dict[str, DetSet], old_size, new_size, flip: str, thresh=None) -> Dict[str, DetSet]:
    """BASE_Trainer.preprocess_results, base.py:128-136: 'RPN_AUG' (when collected) replaces 'RPN'."""
    out = {"RCNN": process(results["RCNN"], old_size, new_size, flip, thresh)}
    out["RPN"] = process(results["RPN_AUG" if "RPN_AUG" in results else "RPN"], old_size, new_size, flip, thresh)
    return out


def resize_boxes(boxes: torch.Tensor, size) -> torch.Tensor:
    """GDINO.resize_boxes, gdino.py:144-160: cxcywh in [0,1] -> xyxy in pixels (per box: scale, x1y1 = c - wh/2,
    x2y2 = wh + x1y1). An empty input is returned as is."""
    h, w = size
    if boxes.shape[0] == 0:
        return boxes
    b = boxes * torch.tensor([w, h, w, h], dtype=torch.float32)
    xy = b[:, :2] - b[:, 2:] / 2
    return torch.cat((xy, b[:, 2:] + xy), dim=1)


def gdino_collect(ori: DetSet, method: str, rcnn_thresh: float, rpn_thresh: float, nms_thresh: float,
                  aug: Optional[DetSet] = None):
    """GDINO_PROCESSOR.post_process without ZOOM (gdino_processor.py:287-298) + .nms (:164-182): score
    thresholds for the two tags, then MyNMS(method).nms per tag; with an AUG set, 'RPN_AUG' = the NMS of the NMS'ed
    RPN set followed by the AUG detections (:295-297)."""
    out = {}
    for tag, thr in (("RCNN", rcnn_thresh), ("RPN", rpn_thresh)):
        keep = ori["scores"] >= thr
        sub = {k: v[keep] for k, v in ori.items()}
        _, b, s, p, l = mynms(method, sub["pred_boxes"], sub["scores"], sub["probs"], sub["pred_classes"], nms_thresh)
        out[tag] = {"pred_boxes": b, "scores": s, "pred_classes": l, "probs": p}
    if aug is not None:
        both = {k: torch.cat((out["RPN"][k], aug[k])) for k in ("pred_boxes", "scores", "pred_classes", "probs")}
        _, b, s, p, l = mynms(method, both["pred_boxes"], both["scores"], both["probs"], both["pred_classes"], nms_thresh)
        out["RPN_AUG"] = {"pred_boxes": b, "scores": s, "pred_classes": l, "probs": p}
    return out


def rpn_distillation_loss(pred_logits: torch.Tensor, distillation_labels: torch.Tensor, teacher_probs: torch.Tensor,
                          weight: float = 1.0):
    """DualTeacherRPN.losses(only_distillation=True), rpn.py:326-340: KLDivLoss(reduction='mean') between
    log([p, 1-p] + 1e-7), p = sigmoid(logit), and [q, 1-q] over the anchors with distillation label > 0.
    Returns None when no anchor is selected (the reference then reports no such loss)."""
    valid = distillation_labels > 0
    p = torch.sigmoid(pred_logits[valid])
    p = torch.stack((p, 1 - p), dim=1)
    q = teacher_probs[valid]
    q = torch.stack((q, 1 - q), dim=1)
    if valid.float().sum() == 0:
        return None
    loss = torch.nn.functional.kl_div(torch.log(p + 1e-7), q, reduction="mean") * weight
    assert not torch.isnan(loss)
    return loss


def roi_distillation_loss(scores_c: torch.Tensor, gt_probs: torch.Tensor, weight: float = 1.0):
    """FastRCNNOutputLayers.losses, fast_rcnn.py:541-545: KLDivLoss(reduction='mean')(log(softmax(scores_c) + 1e-7),
    gt_probs) over the private (C) boxes' class logits."""
    return torch.nn.functional.kl_div(torch.log(torch.softmax(scores_c, dim=1) + 1e-7), gt_probs, reduction="mean") * weight


# ------------------------------------------------------------------------------------------------
# A12  coin/layers/nms.py  probabilistic-fusion NMS (MyNMS)
# ------------------------------------------------------------------------------------------------
def decode_method(method: str) -> Tuple[Optional[str], Optional[str]]:
    """nms.py:61-82: two letters -> (score_method, box_method); 'mm' and 'nms' mean plain NMS."""
    if method == "nms":
        return None, None
    assert len(method) == 2
    sm = {"p": "probEn", "a": "avg", "m": "max"}[method[0]]
    bm = {"s": "s-avg", "a": "avg", "m": "max"}[method[1]]
    if sm == "max" and bm == "max":
        return None, None
    return sm, bm


def _fusion_nms_core(nms_boxes, boxes, probs, labels, thr, score_method, box_method):
    """nms.py:84-194 on one offset-or-class-restricted set."""
    n = boxes.shape[0]
    x1, y1, x2, y2 = nms_boxes.unbind(1)
    area = (x2 - x1 + 1) * (y2 - y1 + 1)
    score = probs[torch.arange(n), labels]
    alive = score.argsort(descending=True)
    rows = []
    while alive.numel() > 0:
        i, rest = alive[0], alive[1:]
        w = torch.clamp(torch.min(x2[i], x2[rest]) - torch.max(x1[i], x1[rest]) + 1, min=0.0)
        h = torch.clamp(torch.min(y2[i], y2[rest]) - torch.max(y1[i], y1[rest]) + 1, min=0.0)
        inter = w * h
        ovr = inter / (area[i] + area[rest] - inter)
        hit = ovr > thr
        members = torch.cat((rest[hit], i.view(1)))  # matched ones first, the pivot last (:130-133)
        if members.numel() > 1:
            m_prob, m_score, m_box, m_lab = probs[members], score[members], boxes[members], labels[members]
            cls = torch.unique(m_lab)
            assert cls.numel() == 1
            if score_method == "probEn":
                assert bool((m_prob.max(1)[1] == m_lab).all())
                e = torch.exp(torch.log(m_prob).sum(dim=0))
                f_prob = e / e.sum()
                f_score = f_prob[cls[0]]
            elif score_method == "avg":
                f_prob, f_score = m_prob.mean(dim=0), m_score.mean()
            else:  # max
                top = torch.argmax(m_score)
                f_prob, f_score = m_prob[top], m_score[top]
            if box_method == "s-avg":
                wgt = m_score / m_score.sum()
                f_box = (m_box * wgt[:, None]).sum(dim=0)
            elif box_method == "avg":
                f_box = m_box.sum(dim=0) / members.numel()
            else:
                f_box = m_box[torch.argmax(m_score)]
            rows.append((i, f_box, f_score, f_prob, cls[0]))
        else:
            rows.append((i, boxes[i], score[i], probs[i], labels[i]))
        alive = rest[~hit]
    keep = torch.stack([r[0] for r in rows])
    o_box = torch.stack([r[1] for r in rows])
    o_score = torch.stack([r[2] for r in rows])
    o_prob = torch.stack([r[3] for r in rows])
    o_cls = torch.stack([r[4] for r in rows])
    order = o_score.argsort(descending=True)
    return keep[order], o_box[order], o_score[order], o_prob[order], o_cls[order]


def mynms(method: str, boxes, scores, probs, idxs, thr):
    """MyNMS(method).nms(boxes, scores, probs, idxs, thr)  nms.py:205-238."""
    sm, bm = decode_method(method)
    if sm is None:
        keep = d2_ref.batched_nms(boxes, scores, idxs, thr)
        return keep, boxes[keep], scores[keep], probs[keep], idxs[keep]
    assert boxes.shape[-1] == 4
    if len(boxes) < 40000:
        boxes = boxes.float()
        if boxes.numel() == 0:
            return torch.empty((0,), dtype=torch.int64), boxes, scores, probs, idxs
        off = idxs.to(boxes) * (boxes.max() + torch.tensor(1).to(boxes))
        return _fusion_nms_core(boxes + off[:, None], boxes, probs, idxs, thr, sm, bm)
    kept = torch.zeros_like(scores, dtype=torch.bool)
    parts = []
    for cid in torch.unique(idxs).tolist():
        sel = (idxs == cid).nonzero().view(-1)
        k, b, s, p, l = _fusion_nms_core(boxes[sel], boxes[sel], probs[sel], idxs[sel], thr, sm, bm)
        parts.append((b, s, p, l))
        kept[sel[k]] = True
    b, s, p, l = (torch.cat([q[i] for q in parts], dim=0) for i in range(4))
    order = s.argsort(descending=True)
    return kept.nonzero().view(-1)[order], b[order], s[order], p[order], l[order]


def weighted_box_fusion_split(box_a, box_b, score_a, score_b):
    """nms.py:24-31 (weights first, then multiply, then add)."""
    s = torch.stack((score_a, score_b), dim=1)
    w = s / s.sum(dim=1, keepdim=True)
    return box_a * w[:, 0:1] + box_b * w[:, 1:]


# ------------------------------------------------------------------------------------------------
# A11  coin/modeling/roi_heads/fast_rcnn.py:116-175
# ------------------------------------------------------------------------------------------------
def fast_rcnn_inference_single_image(boxes, scores, image_shape, score_thresh, nms_thresh, topk):
    valid = torch.isfinite(boxes).all(dim=1) & torch.isfinite(scores).all(dim=1)
    if not valid.all():
        boxes, scores = boxes[valid], scores[valid]
    probs = scores.clone()
    scores = scores[:, :-1]
    kreg = boxes.shape[1] // 4
    boxes = d2_ref.box_clip(boxes.reshape(-1, 4), image_shape).view(-1, kreg, 4)
    fmask = scores > score_thresh
    finds = fmask.nonzero()
    boxes = boxes[finds[:, 0], 0] if kreg == 1 else boxes[fmask]
    scores = scores[fmask]
    probs = probs[finds[:, 0]]
    keep = d2_ref.batched_nms(boxes, scores, finds[:, 1], nms_thresh)
    if topk >= 0:
        keep = keep[:topk]
    finds = finds[keep]
    return {"pred_boxes": boxes[keep], "scores": scores[keep], "probs": probs[keep],
            "pred_classes": finds[:, 1]}, finds[:, 0]


# ------------------------------------------------------------------------------------------------
# A8  relabelling epilogues of the fused IoU+Matcher step
# ------------------------------------------------------------------------------------------------
def relabel_roi(matched_idxs, matched_labels, len_a, len_b, len_c):
    """clip_roi_heads.py:358-362: fg matches that fell on a private (C) box are ignored (-1)."""
    lab = matched_labels.clone()
    in_c = (matched_idxs >= len_a + len_b) & (matched_idxs < len_a + len_b + len_c)
    lab[in_c & ~(matched_labels == 0)] = -1
    return lab


def relabel_rpn(matched_idxs, labels, len_a, len_c):
    """rpn.py:214-228: returns (gt_labels, matched_idxs, distillation_idxs, distillation_labels)."""
    idx, lab = matched_idxs.clone(), labels.clone()
    in_c = (idx >= len_a) & (idx < len_a + len_c)
    fg_c = in_c & ~(lab == 0)
    lab[fg_c] = -1
    dist_idx = matched_idxs - len_a
    dist_idx[~fg_c] = 0
    idx[in_c] = 0
    dist_lab = lab.clone()
    dist_lab[fg_c] = 1
    dist_lab[~fg_c] = 0
    return lab, idx, dist_idx, dist_lab


def relabel_pretrain(matched_idxs, labels, n_gt, n_no_thresh):
    """The 'pre_train' branches (clip_roi_heads.py:305-309, rpn.py:161-165): a match on a `no_thresh_boxes` row is
    neither foreground nor (unless the Matcher already said background) kept: label -1, and its index is reset to 0.
    Returns (labels, matched_idxs)."""
    idx, lab = matched_idxs.clone(), labels.clone()
    in_nt = (idx >= n_gt) & (idx < n_gt + n_no_thresh)
    lab[in_nt & ~(lab == 0)] = -1
    idx[in_nt] = 0
    return lab, idx


# ------------------------------------------------------------------------------------------------
# A9  knowledge separation: coin/utils/util.py:434-507 + coin/engine/trainer.py:338-485
# ------------------------------------------------------------------------------------------------
def delete_duplicate_boxes(d: DetSet, return_split: bool = False,
                           choose: Callable[[int], int] = first_choice):
    """util.py:434-457. Groups rows by the fp32 sum of their 4 coordinates; a group is a duplicate
    set only when the summed difference to its first row is exactly zero (the reference's weak
    test, kept as is)."""
    boxes = d["gt_boxes"]
    key = boxes.sum(1)
    uniq = torch.unique(key)
    member = torch.eq(uniq.unsqueeze(1), key)
    member = member[member.sum(1) != 1]
    groups = []
    for g in range(member.size(0)):
        rows = member[g]
        if (boxes[rows] - boxes[rows][0]).sum() == 0:
            grp = take(d, rows)
            groups.append(grp if return_split else take(grp, choose(length(grp))))
        else:
            member[g][rows.nonzero()[:, 0]] = False
    singles = take(d, (member.sum(0) == 0).nonzero()[:, 0])
    if return_split:
        return singles, groups
    return cat([singles] + groups)


def self_clusters(boxes: torch.Tensor, thresh: float, set_order: str = "cpython") -> List[List[int]]:
    """util.py:459-482 (filter_result + find_same): index clusters of size != 1 among boxes whose
    mutual IoU >= thresh, closed transitively by the reference's recursive set union."""
    adj = d2_ref.pairwise_iou(boxes, boxes) >= thresh
    sets = [set(adj[i].nonzero()[:, 0].tolist()) for i in range(boxes.shape[0])]

    def absorb(path, i):
        for j in _ordered(sets[i], set_order):
            if j != i and j not in path:
                if sets[j] - sets[i]:
                    sets[i] = sets[i] | absorb(path + [i], j)
        return sets[i]

    for i in range(len(sets)):
        for j in _ordered(sets[i], set_order):
            if j != i:
                sets[i] = sets[i] | absorb([i], j)
        for j in sets[i]:
            if j != i:
                sets[j] = set()
    return [_ordered(s, set_order) for s in sets if len(s) > 1]


def online_boxes_merging(online: DetSet, common_off: DetSet, common_on: DetSet,
                         set_order: str = "cpython"):
    """util.py:484-507: resolve cloud boxes that overlap each other at IoU >= 0.95 with
    different classes."""
    for cluster in self_clusters(online["gt_boxes"], 0.95, set_order):
        grp = take(online, torch.tensor(cluster, dtype=torch.int64))
        assert grp["gt_classes"].unique().size(0) != 1
        hit = torch.eq(grp["gt_boxes"].unsqueeze(1), common_on["gt_boxes"]).sum(-1) == 4
        touched = torch.unique(hit.nonzero()[:, 1])
        untouched_mask = torch.ones(length(common_on))
        untouched_mask[touched] = 0
        untouched = untouched_mask.nonzero()[:, 0]
        with_first = hit[0].nonzero()[:, 0]
        clip_cls = common_off["gt_classes"][with_first].unique()
        if clip_cls.size(0) == 1:
            agree = common_on["gt_classes"][touched] == clip_cls
            if agree.sum() != 0:
                touched = touched[agree]
        else:
            differ = common_on["gt_classes"][touched] != common_off["gt_classes"][touched]
            touched = touched[differ]
        common_on = cat([take(common_on, untouched), take(common_on, touched)])
        common_off = cat([take(common_off, untouched), take(common_off, touched)])
    return common_off, common_on


def match_dual_teacher(online: DetSet, offline: DetSet, tag: str, iou_thr: float = 0.5,
                       weight_for_box_a: float = 1.0,
                       choose: Callable[[int], int] = first_choice, set_order: str = "cpython"):
    """trainer.py:338-461. ``online`` = cloud detections (after process()), ``offline`` =
    CLIP-detector detections; both with fields gt_boxes, gt_classes, scores, probs.
    Returns (A, B or None, C) as dicts."""
    nc, nd = length(online), length(offline)
    if nc == 0 and nd == 0:
        com_on, com_off, off_only, on_only = online, offline, [offline], online
    elif nc == 0:
        fg = offline["scores"] > 0.8
        com_on = com_off = take(offline, fg)
        off_only, on_only = [take(offline, ~fg)], online
    elif nd == 0:
        com_on = com_off = online
        off_only, on_only = [offline], offline
    else:
        uniq, dup_groups = delete_duplicate_boxes(offline, return_split=True)
        pairs = (d2_ref.pairwise_iou(online["gt_boxes"], uniq["gt_boxes"]) >= iou_thr).nonzero()
        on_parts, off_parts = [take(online, pairs[:, 0])], [take(uniq, pairs[:, 1])]
        left = set(range(length(uniq))) - set(pairs[:, 1].tolist())
        off_only = [take(uniq, torch.tensor(_ordered(left, set_order), dtype=torch.int64))]
        used_on = pairs[:, 0].tolist()
        for grp in dup_groups:
            gp = (d2_ref.pairwise_iou(online["gt_boxes"], grp["gt_boxes"]) >= iou_thr).nonzero()
            if gp.size(0) != 0:
                i0 = int(gp[0, 0])
                same = grp["gt_classes"] == online["gt_classes"][i0]
                on_parts.append(take(online, i0))
                used_on.append(i0)
                off_parts.append(take(grp, same) if same.sum() >= 1
                                 else take(grp, choose(length(grp))))
            else:
                off_only.append(take(grp, choose(length(grp))))
        com_off, com_on = online_boxes_merging(online, cat(off_parts), cat(on_parts), set_order)
        rest = set(range(nc)) - set(used_on)
        on_only = take(online, torch.tensor(_ordered(rest, set_order), dtype=torch.int64))

    c = cat(off_only + [on_only])
    C = {"gt_boxes": c["gt_boxes"], "gt_classes": c["gt_classes"], "gt_scores": c["scores"],
         "gt_probs": c["probs"]}

    def merged(on_s: DetSet, off_s: DetSet):
        if weight_for_box_a != 1.0:
            return weighted_box_fusion_split(on_s["gt_boxes"], off_s["gt_boxes"], on_s["scores"],
                                             off_s["scores"])
        return on_s["gt_boxes"]

    def pack(on_s: DetSet, off_s: DetSet, split_classes: bool) -> DetSet:
        out = {"gt_boxes": merged(on_s, off_s)}
        if split_classes:
            out["gt_classes_offline"], out["gt_classes_online"] = off_s["gt_classes"], on_s["gt_classes"]
        else:
            out["gt_classes"] = off_s["gt_classes"]
        out["gt_scores_online"], out["gt_scores_offline"] = on_s["scores"], off_s["scores"]
        out["gt_probs_online"], out["gt_probs_offline"] = on_s["probs"], off_s["probs"]
        # Instances.set asserts equal field lengths (trainer.py:404,440): a duplicate group with several members of
        # the matched class appends several offline rows against one online row (trainer.py:379-383) and fails here
        assert len({int(v.shape[0]) for v in out.values()}) == 1, "Adding a field of another length to an Instances"
        return delete_duplicate_boxes(out, choose=choose)

    if tag == "RCNN":
        same = com_off["gt_classes"] == com_on["gt_classes"]
        A = pack(take(com_on, same), take(com_off, same), False)
        B = pack(take(com_on, ~same), take(com_off, ~same), True)
        clash = torch.eq(B["gt_boxes"].unsqueeze(1), A["gt_boxes"]).sum(-1) == 4
        B = take(B, clash.sum(1) == 0)
    elif tag == "RPN":
        A, B = pack(com_on, com_off, False), None
    else:
        raise ValueError(tag)
    return A, B, C


def box_reg_loss(weights, proposal_boxes, gt_boxes, pred_deltas, gt_classes, num_classes: int, smooth_l1_beta: float = 0.0,
                 normalizer=None):
    """coin/modeling/roi_heads/fast_rcnn.py:601-646 (box_reg_loss_type == "smooth_l1") with fvcore's smooth_l1_loss
    (beta < 1e-5 -> L1), reduction "sum", normalised by the number of regions (or `normalizer`)."""
    from . import d2_ref
    box_dim = proposal_boxes.shape[1]
    fg_inds = torch.nonzero((gt_classes >= 0) & (gt_classes < num_classes), as_tuple=True)[0]
    if pred_deltas.shape[1] == box_dim:
        fg_pred_deltas = pred_deltas[fg_inds]
    else:
        fg_pred_deltas = pred_deltas.view(-1, num_classes, box_dim)[fg_inds, gt_classes[fg_inds]]
    gt_pred_deltas = d2_ref.Box2BoxTransform(weights).get_deltas(proposal_boxes[fg_inds], gt_boxes[fg_inds])
    n = torch.abs(fg_pred_deltas - gt_pred_deltas)
    if smooth_l1_beta < 1e-5:
        loss = n.sum()
    else:
        loss = torch.where(n < smooth_l1_beta, 0.5 * n ** 2 / smooth_l1_beta, n - 0.5 * smooth_l1_beta).sum()
    if normalizer is not None:
        return loss / normalizer
    return loss / max(gt_classes.numel(), 1.0)


# ------------------------------------------------------------------------------------------------
# SURVEY 8(f) rank 4  coin/evaluation/cloud_pascal_voc_evaluation.py:173-319  VOC AP for one class
# ------------------------------------------------------------------------------------------------
def voc_ap(rec, prec, use_07_metric: bool = False) -> float:
    """cloud_pascal_voc_evaluation.py:173-202."""
    import numpy as np
    if use_07_metric:
        ap = 0.0
        for t in np.arange(0.0, 1.1, 0.1):
            p = 0 if np.sum(rec >= t) == 0 else np.max(prec[rec >= t])
            ap = ap + p / 11.0
        return float(ap)
    mrec = np.concatenate(([0.0], rec, [1.0]))
    mpre = np.concatenate(([0.0], prec, [0.0]))
    for i in range(mpre.size - 1, 0, -1):
        mpre[i - 1] = np.maximum(mpre[i - 1], mpre[i])
    i = np.where(mrec[1:] != mrec[:-1])[0]
    return float(np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1]))


def voc_eval_class(det_image, det_conf, det_boxes, gt_boxes_per_image, gt_difficult_per_image, ovthresh: float = 0.5,
                   use_07_metric: bool = False, order=None):
    """voc_eval (cloud_pascal_voc_evaluation.py:205-319) for one class on arrays instead of files: det_image[i] indexes the
    per-image ground-truth lists; legacy '+1' IoU in float64; a detection is a true positive when its best ground-truth box
    (first maximum) overlaps it by more than ovthresh, is not 'difficult' and has not been claimed by a more confident one.
    order: the np.argsort(-confidence) of the original (unstable on ties); None: stable (lower index first)."""
    import numpy as np
    det_conf = np.asarray(det_conf, dtype=np.float64)
    bb_all = np.asarray(det_boxes, dtype=np.float64).reshape(-1, 4)
    npos = int(sum(int((~np.asarray(d, dtype=bool)).sum()) for d in gt_difficult_per_image))
    claimed = [np.zeros(len(d), dtype=bool) for d in gt_difficult_per_image]
    sorted_ind = np.argsort(-det_conf, kind="stable") if order is None else np.asarray(order)
    nd = len(sorted_ind)
    tp, fp = np.zeros(nd), np.zeros(nd)
    for d in range(nd):
        img = int(det_image[sorted_ind[d]])
        bb = bb_all[sorted_ind[d]]
        bbgt = np.asarray(gt_boxes_per_image[img], dtype=np.float64).reshape(-1, 4)
        ovmax, jmax = -np.inf, -1
        if bbgt.size > 0:
            iw = np.maximum(np.minimum(bbgt[:, 2], bb[2]) - np.maximum(bbgt[:, 0], bb[0]) + 1.0, 0.0)
            ih = np.maximum(np.minimum(bbgt[:, 3], bb[3]) - np.maximum(bbgt[:, 1], bb[1]) + 1.0, 0.0)
            inters = iw * ih
            uni = (bb[2] - bb[0] + 1.0) * (bb[3] - bb[1] + 1.0) + (bbgt[:, 2] - bbgt[:, 0] + 1.0) * (bbgt[:, 3] - bbgt[:, 1] + 1.0) - inters
            overlaps = inters / uni
            ovmax, jmax = np.max(overlaps), int(np.argmax(overlaps))
        if ovmax > ovthresh:
            if not gt_difficult_per_image[img][jmax]:
                if not claimed[img][jmax]:
                    tp[d] = 1.0
                    claimed[img][jmax] = True
                else:
                    fp[d] = 1.0
        else:
            fp[d] = 1.0
    fp, tp = np.cumsum(fp), np.cumsum(tp)
    rec = tp / float(npos)
    prec = tp / np.maximum(tp + fp, np.finfo(np.float64).eps)
    return rec, prec, voc_ap(rec, prec, use_07_metric)

"""CPU restatement of the detectron2-0.5 / torchvision operators on the COIN RoI path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py). All tensors are CPU tensors.

detectron2 0.5 is pinned by the reference in docs/Environment.md:51-52 but is neither vendored
under /root/reference nor installable offline, so each function restates the published algorithm
and cites the reference call site it serves ("[d2]" = detectron2 0.5 module it restates).
roi_align / nms are executed through the torchvision CPU operators of this image (the same
dependency the reference reaches, at a newer version) and cross-checked by oracle/scalar_ref.c.
"""
import math
from typing import List, Sequence, Tuple

import torch
import torchvision

_SCALE_CLAMP = math.log(1000.0 / 16)


# ------------------------------------------------------------------------------------------------
# Boxes helpers  [d2] structures/boxes.py -- used at fast_rcnn.py:145-147, base.py:89, gdino.py:136
# ------------------------------------------------------------------------------------------------
def box_area(b: torch.Tensor) -> torch.Tensor:
    return (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])


def box_clip(b: torch.Tensor, image_size: Tuple[int, int]) -> torch.Tensor:
    """Boxes.clip((h, w)): x to [0, w], y to [0, h]; returns a new tensor."""
    h, w = image_size
    x1 = b[:, 0].clamp(min=0, max=w)
    y1 = b[:, 1].clamp(min=0, max=h)
    x2 = b[:, 2].clamp(min=0, max=w)
    y2 = b[:, 3].clamp(min=0, max=h)
    return torch.stack((x1, y1, x2, y2), dim=-1)


def box_scale(b: torch.Tensor, sx: float, sy: float) -> torch.Tensor:
    """Boxes.scale(scale_x, scale_y) (in place in d2; functional here)."""
    out = b.clone()
    out[:, 0::2] *= sx
    out[:, 1::2] *= sy
    return out


def box_nonempty(b: torch.Tensor, threshold: float = 0.0) -> torch.Tensor:
    return ((b[:, 2] - b[:, 0]) > threshold) & ((b[:, 3] - b[:, 1]) > threshold)


# ------------------------------------------------------------------------------------------------
# pairwise_iou  [d2] structures/boxes.py -- trainer.py:364,373; util.py:468;
#               clip_roi_heads.py:301,311,353; rpn.py:159,169,212
# ------------------------------------------------------------------------------------------------
def pairwise_iou(b1: torch.Tensor, b2: torch.Tensor) -> torch.Tensor:
    b1 = b1.reshape(-1, 4).float()
    b2 = b2.reshape(-1, 4).float()
    a1, a2 = box_area(b1), box_area(b2)
    wh = torch.min(b1[:, None, 2:], b2[:, 2:]) - torch.max(b1[:, None, :2], b2[:, :2])
    wh.clamp_(min=0)
    inter = wh.prod(dim=2)
    return torch.where(inter > 0, inter / (a1[:, None] + a2 - inter), torch.zeros(1, dtype=inter.dtype))


# ------------------------------------------------------------------------------------------------
# Matcher  [d2] modeling/matcher.py -- clip_roi_heads.py:126-130,304,314,356; rpn.py:160,170,213
# ------------------------------------------------------------------------------------------------
class Matcher:
    def __init__(self, thresholds: Sequence[float], labels: Sequence[int],
                 allow_low_quality_matches: bool = False):
        thresholds = list(thresholds)
        assert thresholds[0] > 0
        thresholds.insert(0, -float("inf"))
        thresholds.append(float("inf"))
        assert all(lo <= hi for lo, hi in zip(thresholds[:-1], thresholds[1:]))
        assert all(l in (-1, 0, 1) for l in labels)
        assert len(labels) == len(thresholds) - 1
        self.thresholds, self.labels = thresholds, list(labels)
        self.allow_low_quality_matches = allow_low_quality_matches

    def __call__(self, q: torch.Tensor):
        assert q.dim() == 2
        if q.numel() == 0:
            return (q.new_full((q.size(1),), 0, dtype=torch.int64),
                    q.new_full((q.size(1),), self.labels[0], dtype=torch.int8))
        assert torch.all(q >= 0)
        vals, matches = q.max(dim=0)
        lab = matches.new_full(matches.size(), 1, dtype=torch.int8)
        for l, lo, hi in zip(self.labels, self.thresholds[:-1], self.thresholds[1:]):
            lab[(vals >= lo) & (vals < hi)] = l
        if self.allow_low_quality_matches:
            best_per_gt = q.max(dim=1).values
            cols = torch.nonzero(q == best_per_gt[:, None], as_tuple=True)[1]
            lab[cols] = 1
        return matches, lab


# ------------------------------------------------------------------------------------------------
# Box2BoxTransform  [d2] modeling/box_regression.py -- fast_rcnn.py:297,619-622,691,729; d2 RPN
# ------------------------------------------------------------------------------------------------
class Box2BoxTransform:
    def __init__(self, weights, scale_clamp: float = _SCALE_CLAMP):
        self.weights, self.scale_clamp = tuple(weights), scale_clamp

    def get_deltas(self, src: torch.Tensor, tgt: torch.Tensor) -> torch.Tensor:
        sw, sh = src[:, 2] - src[:, 0], src[:, 3] - src[:, 1]
        scx, scy = src[:, 0] + 0.5 * sw, src[:, 1] + 0.5 * sh
        tw, th = tgt[:, 2] - tgt[:, 0], tgt[:, 3] - tgt[:, 1]
        tcx, tcy = tgt[:, 0] + 0.5 * tw, tgt[:, 1] + 0.5 * th
        wx, wy, ww, wh = self.weights
        dx = wx * (tcx - scx) / sw
        dy = wy * (tcy - scy) / sh
        dw = ww * torch.log(tw / sw)
        dh = wh * torch.log(th / sh)
        assert (sw > 0).all().item(), "Input boxes to Box2BoxTransform are not valid!"
        return torch.stack((dx, dy, dw, dh), dim=1)

    def apply_deltas(self, deltas: torch.Tensor, boxes: torch.Tensor) -> torch.Tensor:
        deltas = deltas.float()
        boxes = boxes.to(deltas.dtype)
        w, h = boxes[:, 2] - boxes[:, 0], boxes[:, 3] - boxes[:, 1]
        cx, cy = boxes[:, 0] + 0.5 * w, boxes[:, 1] + 0.5 * h
        wx, wy, ww, wh = self.weights
        dx, dy = deltas[:, 0::4] / wx, deltas[:, 1::4] / wy
        dw, dh = deltas[:, 2::4] / ww, deltas[:, 3::4] / wh
        dw = torch.clamp(dw, max=self.scale_clamp)
        dh = torch.clamp(dh, max=self.scale_clamp)
        pcx = dx * w[:, None] + cx[:, None]
        pcy = dy * h[:, None] + cy[:, None]
        pw = torch.exp(dw) * w[:, None]
        ph = torch.exp(dh) * h[:, None]
        out = torch.stack((pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph), dim=-1)
        return out.reshape(deltas.shape)


# ------------------------------------------------------------------------------------------------
# ROIAlign / ROIPooler  [d2] layers/roi_align.py, modeling/poolers.py -- clip_roi_heads.py:51-62,142-176
# ------------------------------------------------------------------------------------------------
def roi_align(x, rois, output_size, spatial_scale, sampling_ratio, aligned=True):
    """[d2] ROIAlign.forward: rois are cast to the input dtype first; under autocast torchvision
    computes in fp32 and casts the result back to the input dtype (csrc/ops/autocast)."""
    if isinstance(output_size, int):
        output_size = (output_size, output_size)
    rois = rois.to(dtype=x.dtype)
    out = torchvision.ops.roi_align(x.float(), rois.float(), output_size, spatial_scale,
                                    sampling_ratio, aligned)
    return out.to(x.dtype)


def assign_boxes_to_levels(box_lists: List[torch.Tensor], min_level: int, max_level: int,
                           canonical_box_size: int = 224, canonical_level: int = 4) -> torch.Tensor:
    sizes = torch.sqrt(torch.cat([box_area(b) for b in box_lists]))
    lvl = torch.floor(canonical_level + torch.log2(sizes / canonical_box_size + 1e-8))
    lvl = torch.clamp(lvl, min=min_level, max=max_level)
    return lvl.to(torch.int64) - min_level


def pooler_format(box_lists: List[torch.Tensor]) -> torch.Tensor:
    parts = []
    for i, b in enumerate(box_lists):
        parts.append(torch.cat((torch.full((len(b), 1), i, dtype=b.dtype), b), dim=1))
    return torch.cat(parts, dim=0) if parts else torch.zeros((0, 5))


def roi_pooler(x: List[torch.Tensor], box_lists: List[torch.Tensor], output_size, scales,
               sampling_ratio=0, pooler_type="ROIAlignV2", canonical_box_size=224,
               canonical_level=4) -> torch.Tensor:
    if isinstance(output_size, int):
        output_size = (output_size, output_size)
    aligned = {"ROIAlign": False, "ROIAlignV2": True}[pooler_type]
    min_level, max_level = -math.log2(scales[0]), -math.log2(scales[-1])
    assert math.isclose(min_level, int(min_level)) and math.isclose(max_level, int(max_level))
    min_level, max_level = int(min_level), int(max_level)
    assert len(scales) == max_level - min_level + 1 and len(x) == len(scales)
    c = x[0].shape[1]
    if sum(len(b) for b in box_lists) == 0:
        return torch.zeros((0, c) + tuple(output_size), dtype=x[0].dtype)
    rois = pooler_format(box_lists)
    if len(scales) == 1:
        return roi_align(x[0], rois, output_size, scales[0], sampling_ratio, aligned)
    lvl = assign_boxes_to_levels(box_lists, min_level, max_level, canonical_box_size, canonical_level)
    out = torch.zeros((rois.shape[0], c) + tuple(output_size), dtype=x[0].dtype)
    for l, (xl, s) in enumerate(zip(x, scales)):
        inds = torch.nonzero(lvl == l, as_tuple=True)[0]
        out.index_put_((inds,), roi_align(xl, rois[inds], output_size, s, sampling_ratio, aligned))
    return out


# ------------------------------------------------------------------------------------------------
# nms / batched_nms  [tv] ops/boxes.py + [d2] layers/nms.py -- fast_rcnn.py:164; nms.py:207;
#                    clip_rcnn.py:161; d2 find_top_rpn_proposals (<- rpn.py:113)
# ------------------------------------------------------------------------------------------------
def nms(boxes, scores, thr):
    return torchvision.ops.nms(boxes.float(), scores, thr)


def batched_nms_strategy(n_boxes: int) -> str:
    """Which torchvision strategy the CPU path of the reference's dependency picks for this size."""
    return "vanilla" if n_boxes * 4 > 4000 else "trick"


def batched_nms(boxes, scores, idxs, thr):
    """[d2] layers/nms.py::batched_nms wrapper over torchvision.ops.boxes.batched_nms."""
    assert boxes.shape[-1] == 4
    if len(boxes) < 40000:
        return torchvision.ops.boxes.batched_nms(boxes.float(), scores, idxs, thr)
    result_mask = scores.new_zeros(scores.size(), dtype=torch.bool)
    for cid in torch.unique(idxs).tolist():
        mask = (idxs == cid).nonzero().view(-1)
        keep = nms(boxes[mask], scores[mask], thr)
        result_mask[mask[keep]] = True
    keep = result_mask.nonzero().view(-1)
    return keep[scores[keep].argsort(descending=True)]


# ------------------------------------------------------------------------------------------------
# d2 RPN helpers used by the "next" rows (anchor generation and top-k proposal selection)
# ------------------------------------------------------------------------------------------------
def cell_anchors(sizes=(32, 64, 128, 256, 512), ratios=(0.5, 1.0, 2.0)) -> torch.Tensor:
    out = []
    for s in sizes:
        area = s ** 2.0
        for r in ratios:
            w = math.sqrt(area / r)
            h = r * w
            out.append([-w / 2.0, -h / 2.0, w / 2.0, h / 2.0])
    return torch.tensor(out, dtype=torch.float32)


def grid_anchors(hf: int, wf: int, stride: int, base: torch.Tensor, offset: float = 0.0) -> torch.Tensor:
    sx = torch.arange(offset * stride, wf * stride, step=stride, dtype=torch.float32)
    sy = torch.arange(offset * stride, hf * stride, step=stride, dtype=torch.float32)
    yy, xx = torch.meshgrid(sy, sx, indexing="ij")
    shifts = torch.stack((xx.reshape(-1), yy.reshape(-1), xx.reshape(-1), yy.reshape(-1)), dim=1)
    return (shifts.view(-1, 1, 4) + base.view(1, -1, 4)).reshape(-1, 4)


def predict_proposals_single(anchors, pred_anchor_deltas, pred_objectness_logits, image_size, nms_thresh: float,
                             pre_nms_topk: int, post_nms_topk: int, min_box_size: float = 0.0, training: bool = False,
                             weights=(1.0, 1.0, 1.0, 1.0)):
    """[d2 0.5] RPN.predict_proposals for ONE image and ONE feature level (<- rpn.py:64,113):
    proposal_generator/rpn.py::_decode_proposals (Box2BoxTransform.apply_deltas on every anchor) followed by
    proposal_utils.py::find_top_rpn_proposals: sort the logits (descending), keep pre_nms_topk, drop non-finite rows
    (FloatingPointError when training), Boxes.clip, Boxes.nonempty(threshold=min_box_size), batched_nms over the
    level ids (one level: plain nms), keep[:post_nms_topk]. Ties of equal logits keep the lower anchor index
    (stable sort; torch's default sort leaves the order of ties unspecified) - the determinism policy of DESIGN.md.
    Returns (proposal_boxes [n,4], objectness_logits [n])."""
    proposals = Box2BoxTransform(weights).apply_deltas(pred_anchor_deltas.float(), anchors.float())
    logits = pred_objectness_logits.reshape(-1).float()
    num = min(logits.numel(), pre_nms_topk)
    sorted_logits, idx = torch.sort(logits, descending=True, stable=True)
    scores, idx = sorted_logits[:num], idx[:num]
    boxes = proposals[idx]
    valid = torch.isfinite(boxes).all(dim=1) & torch.isfinite(scores)
    if not valid.all():
        if training:
            raise FloatingPointError("Predicted boxes or scores contain Inf/NaN. Training has diverged.")
        boxes, scores = boxes[valid], scores[valid]
    boxes = box_clip(boxes, image_size)
    keep = box_nonempty(boxes, threshold=min_box_size)
    if keep.sum().item() != len(boxes):
        boxes, scores = boxes[keep], scores[keep]
    keep = batched_nms(boxes, scores, torch.zeros(len(boxes), dtype=torch.int64), nms_thresh)
    keep = keep[:post_nms_topk]
    return boxes[keep], scores[keep]


def detector_postprocess(boxes: torch.Tensor, image_size, output_height: int, output_width: int):
    """[d2 0.5] modeling/postprocessing.py::detector_postprocess on the box tensor (<- clip_rcnn.py:424 via
    GeneralizedRCNN._postprocess): Boxes.scale(out_w / in_w, out_h / in_h), Boxes.clip((out_h, out_w)), nonempty().
    Returns (boxes of the kept rows, keep mask)."""
    sx, sy = output_width / image_size[1], output_height / image_size[0]
    b = box_clip(box_scale(boxes, sx, sy), (output_height, output_width))
    keep = box_nonempty(b)
    return b[keep], keep


# ------------------------------------------------------------------------------------------------
# detectron2 0.5 modeling/sampling.py::subsample_labels and ROIHeads._sample_proposals
# (<- coin/modeling/roi_heads/clip_roi_heads.py:317,363; coin/modeling/proposal_generator/rpn.py:231)
# ------------------------------------------------------------------------------------------------
def philox_keys(elements, stream_set: int, seed: int, offset: int):
    """Philox4x32-10, first output word, counter = (element, set, offset lo, offset hi), key = (seed lo, seed hi): the
    device generator of coin_subsample_labels restated with numpy integers (the random draw is a policy of the
    replacement, not of the reference, which uses torch.randperm; DESIGN.md 'Determinism policy')."""
    import numpy as np
    e = np.asarray(elements, dtype=np.uint64)
    c0 = e & np.uint64(0xFFFFFFFF)
    c1 = np.full_like(c0, np.uint64(stream_set))
    c2 = np.full_like(c0, np.uint64(offset & 0xFFFFFFFF))
    c3 = np.full_like(c0, np.uint64((offset >> 32) & 0xFFFFFFFF))
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    m32 = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c0
        p1 = np.uint64(0xCD9E8D57) * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & m32, p1 >> np.uint64(32), p1 & m32
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + np.uint64(0x9E3779B9)) & m32
        k1 = (k1 + np.uint64(0xBB67AE85)) & m32
    return c0.astype(np.uint64)


def subsample_labels(labels: torch.Tensor, num_samples: int, positive_fraction: float, bg_label: int, perms=None,
                     seed: int = 0, offset: int = 0):
    """[d2] subsample_labels. perms = (perm1, perm2): the two torch.randperm draws of the original; None: the device
    policy - the num smallest (Philox key, element) pairs of each set, in key order."""
    positive = torch.nonzero((labels != -1) & (labels != bg_label), as_tuple=True)[0]
    negative = torch.nonzero(labels == bg_label, as_tuple=True)[0]
    num_pos = min(positive.numel(), int(num_samples * positive_fraction))
    num_neg = min(negative.numel(), num_samples - num_pos)
    if perms is not None:
        return positive[perms[0][:num_pos]], negative[perms[1][:num_neg]]
    import numpy as np
    out = []
    for s, (cand, num) in enumerate(((positive, num_pos), (negative, num_neg))):
        keys = philox_keys(cand.numpy(), s, seed, offset)
        order = np.lexsort((cand.numpy(), keys))          # by key, ties by element
        out.append(cand[torch.from_numpy(order[:num].copy())])
    return out[0], out[1]


def sample_proposals(matched_idxs, matched_labels, gt_classes, num_classes: int, batch_size_per_image: int,
                     positive_fraction: float, perms=None, seed: int = 0, offset: int = 0):
    """[d2] ROIHeads._sample_proposals -> (sampled_idxs, gt_classes[sampled_idxs])."""
    if gt_classes.numel() > 0:
        gt_classes = gt_classes[matched_idxs]
        gt_classes[matched_labels == 0] = num_classes
        gt_classes[matched_labels == -1] = -1
    else:
        gt_classes = torch.zeros_like(matched_idxs) + num_classes
    fg, bg = subsample_labels(gt_classes, batch_size_per_image, positive_fraction, num_classes, perms, seed, offset)
    sampled = torch.cat([fg, bg], dim=0)
    return sampled, gt_classes[sampled]

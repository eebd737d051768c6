"""The operator API the reference calls, backed by libcoinops (sm_100a CUDA kernels).

Drop-in for the names the COIN reference imports on its RoI path:
    detectron2.layers.ROIAlign, detectron2.modeling.poolers.ROIPooler        (clip_roi_heads.py:13,51-62,142-176)
    detectron2.structures.pairwise_iou                                       (trainer.py:26, clip_roi_heads.py:8, rpn.py:10, util.py:23)
    detectron2.modeling.matcher.Matcher                                      (clip_roi_heads.py:5,126; gdino_processor.py:20,79)
    detectron2.modeling.box_regression.Box2BoxTransform                      (fast_rcnn.py:11,297)
    detectron2.layers.batched_nms                                            (fast_rcnn.py:9,164; nms.py:3,207; clip_rcnn.py:28,161)
    coin.layers.nms.MyNMS / mynms                                            (gdino_processor.py:173-174; train_net.py:86)
Same names, argument meaning and error behaviour; the arithmetic runs in hand-written kernels.
"""
import math
from typing import List, Optional, Sequence, Tuple

import torch
from torch import nn

from . import ops
from .structures import Boxes

_SCALE_CLAMP = math.log(1000.0 / 16)


# ------------------------------------------------------------------------------------------------
# ROIAlign / ROIPooler
# ------------------------------------------------------------------------------------------------
class _ROIAlignFn(torch.autograd.Function):
    """Autograd bridge: forward = coin_roi_align_fwd, backward = coin_roi_align_bwd.
    Multi-level: ``feats`` are the per-level maps, ``levels`` the int32 level of each RoI."""

    @staticmethod
    def forward(ctx, rois, levels, output_size, scales, sampling_ratio, aligned, *feats):
        nhwc = [ops.to_nhwc_f32(f) for f in feats]
        if levels is None:
            # one level (COIN's C4 heads): nothing bounds a RoI's size in feature cells, and a map-sized RoI is a 1-ms CTA for
            # the register-tile kernel - the planned forward pools those few with the separable kernel. One launch computes
            # the launch order (the smallest RoIs go last: scheduling only) and the split.
            perm, plan = ops.roi_launch_plan(rois, scales[0])
            out = ops.roi_align_forward_planned(nhwc, scales, rois, output_size, sampling_ratio, aligned, feats[0].dtype,
                                                plan=plan)
        else:   # ROIPooler assigns large boxes to coarse levels: at most ~28 x 28 cells per RoI by construction
            perm, plan = ops.roi_launch_order(rois), None
            out = ops.roi_align_forward(nhwc, scales, rois, levels, output_size, sampling_ratio, aligned, feats[0].dtype,
                                        perm=perm)
        ctx.save_for_backward(rois, levels if levels is not None else torch.empty(0),
                              perm if perm is not None else torch.empty(0), *(plan if plan is not None else ()))
        ctx.has_levels, ctx.has_perm, ctx.has_plan = levels is not None, perm is not None, plan is not None
        ctx.meta = (output_size, tuple(scales), sampling_ratio, aligned, [tuple(f.shape) for f in feats],
                    [f.dtype for f in feats])
        return out

    @staticmethod
    def backward(ctx, grad_out):
        rois, levels, perm = ctx.saved_tensors[:3]
        plan = tuple(ctx.saved_tensors[3:6]) if ctx.has_plan else None
        output_size, scales, sampling_ratio, aligned, shapes, dtypes = ctx.meta
        grads = ops.roi_align_backward(grad_out, shapes, scales, rois, levels if ctx.has_levels else None, output_size,
                                       sampling_ratio, aligned, dtypes, perm=perm if ctx.has_perm else None, plan=plan)
        return (None, None, None, None, None, None, *grads)


class ROIAlign(nn.Module):
    """detectron2.layers.ROIAlign(output_size, spatial_scale, sampling_ratio, aligned=True)."""

    def __init__(self, output_size, spatial_scale, sampling_ratio, aligned=True):
        super().__init__()
        self.output_size = (output_size, output_size) if isinstance(output_size, int) else tuple(output_size)
        self.spatial_scale = spatial_scale
        self.sampling_ratio = sampling_ratio
        self.aligned = aligned

    def forward(self, input: torch.Tensor, rois: torch.Tensor) -> torch.Tensor:
        assert rois.dim() == 2 and rois.size(1) == 5
        # detectron2 casts the rois to the input dtype before the op; torchvision's autocast wrapper
        # then computes in fp32 and returns the input dtype. Reproduce both casts.
        rois = rois.to(dtype=input.dtype).to(torch.float32)
        return _ROIAlignFn.apply(rois, None, self.output_size, (self.spatial_scale,), self.sampling_ratio,
                                 self.aligned, input)

    def __repr__(self):
        return (f"{self.__class__.__name__}(output_size={self.output_size}, spatial_scale={self.spatial_scale}, "
                f"sampling_ratio={self.sampling_ratio}, aligned={self.aligned})")


def convert_boxes_to_pooler_format(box_lists: List[Boxes]) -> torch.Tensor:
    parts = []
    for i, b in enumerate(box_lists):
        t = b.tensor
        parts.append(torch.cat((torch.full((len(t), 1), i, dtype=t.dtype, device=t.device), t), dim=1))
    return torch.cat(parts, dim=0)


class ROIPooler(nn.Module):
    """detectron2.modeling.poolers.ROIPooler(output_size, scales, sampling_ratio, pooler_type).
    All levels are pooled by ONE kernel launch (per-RoI level lookup), not one launch per level."""

    def __init__(self, output_size, scales, sampling_ratio, pooler_type, canonical_box_size=224, canonical_level=4):
        super().__init__()
        if isinstance(output_size, int):
            output_size = (output_size, output_size)
        assert len(output_size) == 2 and isinstance(output_size[0], int) and isinstance(output_size[1], int)
        self.output_size = tuple(output_size)
        if pooler_type == "ROIAlign":
            self.aligned = False
        elif pooler_type == "ROIAlignV2":
            self.aligned = True
        elif pooler_type in ("ROIPool", "ROIAlignRotated"):
            raise NotImplementedError(f"pooler_type {pooler_type} is outside the COIN hot path (SURVEY.md 8a A1)")
        else:
            raise ValueError("Unknown pooler type: {}".format(pooler_type))
        self.scales = tuple(float(s) for s in scales)
        self.sampling_ratio = sampling_ratio
        min_level, max_level = -math.log2(scales[0]), -math.log2(scales[-1])
        assert math.isclose(min_level, int(min_level)) and math.isclose(max_level, int(max_level)), \
            "Featuremap stride is not power of 2!"
        self.min_level, self.max_level = int(min_level), int(max_level)
        assert len(scales) == self.max_level - self.min_level + 1, "[ROIPooler] Sizes of input featuremaps do not form a pyramid!"
        assert 0 <= self.min_level <= self.max_level
        self.canonical_level = canonical_level
        assert canonical_box_size > 0
        self.canonical_box_size = canonical_box_size

    def forward(self, x: List[torch.Tensor], box_lists: List[Boxes]) -> torch.Tensor:
        num_level_assignments = len(self.scales)
        assert isinstance(x, list) and isinstance(box_lists, list), "Arguments to pooler must be lists"
        assert len(x) == num_level_assignments, \
            "unequal value, num_level_assignments={}, but x is list of {} Tensors".format(num_level_assignments, len(x))
        assert len(box_lists) == x[0].size(0), \
            "unequal value, x[0] batch dim 0 is {}, but box_list has length {}".format(x[0].size(0), len(box_lists))
        if len(box_lists) == 0:
            return torch.zeros((0, x[0].shape[1]) + self.output_size, device=x[0].device, dtype=x[0].dtype)
        rois = convert_boxes_to_pooler_format(box_lists)
        if rois.shape[0] == 0:
            return torch.zeros((0, x[0].shape[1]) + self.output_size, device=x[0].device, dtype=x[0].dtype)
        levels = None
        if num_level_assignments > 1:
            levels = ops.roi_pooler_levels(rois[:, 1:].contiguous(), self.min_level, self.max_level,
                                           self.canonical_box_size, self.canonical_level)
        rois = rois.to(dtype=x[0].dtype).to(torch.float32)
        return _ROIAlignFn.apply(rois, levels, self.output_size, self.scales, self.sampling_ratio, self.aligned, *x)


# ------------------------------------------------------------------------------------------------
# pairwise_iou / Matcher / Box2BoxTransform
# ------------------------------------------------------------------------------------------------
def pairwise_iou(boxes1: Boxes, boxes2: Boxes) -> torch.Tensor:
    """detectron2.structures.pairwise_iou(Boxes[N], Boxes[M]) -> Tensor[N, M]."""
    return ops.pairwise_iou(boxes1.tensor, boxes2.tensor)


class Matcher:
    """detectron2.modeling.matcher.Matcher(thresholds, labels, allow_low_quality_matches)."""

    def __init__(self, thresholds: Sequence[float], labels: Sequence[int], allow_low_quality_matches: bool = False):
        thresholds = list(thresholds)
        assert thresholds[0] > 0
        self._user_thresholds = list(thresholds)
        thresholds.insert(0, -float("inf"))
        thresholds.append(float("inf"))
        assert all(low <= high for (low, high) in zip(thresholds[:-1], thresholds[1:]))
        assert all(l in [-1, 0, 1] for l in labels)
        assert len(labels) == len(thresholds) - 1
        self.thresholds = thresholds
        self.labels = list(labels)
        self.allow_low_quality_matches = allow_low_quality_matches

    def __call__(self, match_quality_matrix: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        assert match_quality_matrix.dim() == 2
        return ops.matcher(match_quality_matrix, self._user_thresholds, self.labels, self.allow_low_quality_matches)

    def match_boxes(self, gt_boxes: Boxes, boxes: Boxes, return_vals: bool = False):
        """Fused ``self(pairwise_iou(gt_boxes, boxes))``: the [N,M] matrix is never materialised."""
        return ops.iou_match(gt_boxes.tensor, boxes.tensor, self._user_thresholds, self.labels,
                             self.allow_low_quality_matches, return_vals)


class Box2BoxTransform:
    """detectron2.modeling.box_regression.Box2BoxTransform(weights, scale_clamp=log(1000/16))."""

    def __init__(self, weights: Tuple[float, float, float, float], scale_clamp: float = _SCALE_CLAMP):
        self.weights = weights
        self.scale_clamp = scale_clamp

    def get_deltas(self, src_boxes: torch.Tensor, target_boxes: torch.Tensor) -> torch.Tensor:
        assert isinstance(src_boxes, torch.Tensor), type(src_boxes)
        assert isinstance(target_boxes, torch.Tensor), type(target_boxes)
        return ops.get_deltas(src_boxes, target_boxes, self.weights)

    def apply_deltas(self, deltas: torch.Tensor, boxes: torch.Tensor,
                     clip_to: Optional[Tuple[float, float]] = None) -> torch.Tensor:
        """``clip_to=(h, w)`` additionally fuses ``Boxes.clip`` (fast_rcnn.py:145-147) into the decode."""
        return ops.apply_deltas(deltas, boxes, self.weights, self.scale_clamp, clip_to)


# ------------------------------------------------------------------------------------------------
# NMS
# ------------------------------------------------------------------------------------------------
def batched_nms(boxes: torch.Tensor, scores: torch.Tensor, idxs: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """detectron2.layers.batched_nms: per-class NMS; kept indices in descending-score order."""
    assert boxes.shape[-1] == 4
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    return ops.batched_nms(boxes.float(), scores, idxs, iou_threshold, "auto")


def nms(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    return ops.nms(boxes.float(), scores, iou_threshold)


class MyNMS:
    """coin.layers.nms.MyNMS: ``method`` is 'nms' or two letters (score: p/a/m, box: s/a/m)."""

    def __init__(self, method):
        self.method = method
        if self.method is not None:
            self.update_cfg()

    def update_cfg(self):
        if self.method != "nms":
            assert len(self.method) == 2
            try:
                self.score_method = {"p": "probEn", "a": "avg", "m": "max"}[self.method[0]]
                self.box_method = {"s": "s-avg", "a": "avg", "m": "max"}[self.method[1]]
            except KeyError:
                raise NotImplementedError
            if self.score_method == "max" and self.box_method == "max":
                self.method = "nms"

    def update(self, method):
        self.method = method
        self.update_cfg()

    def nms(self, boxes: torch.Tensor, scores: torch.Tensor, probs: torch.Tensor, idxs: torch.Tensor,
            iou_threshold: float):
        if self.method == "nms":
            keep = batched_nms(boxes, scores, idxs, iou_threshold)
            return keep, boxes[keep], scores[keep], probs[keep], idxs[keep]
        return self.Probabilistic_Fusion(boxes, scores, probs, idxs, iou_threshold)

    def Probabilistic_Fusion(self, boxes, scores, probs, idxs, iou_threshold):
        assert boxes.shape[-1] == 4
        if boxes.numel() == 0:
            return torch.empty((0,), dtype=torch.int64, device=boxes.device), boxes, scores, probs, idxs
        if len(boxes) < 40000:
            return ops.fusion_nms(boxes.float(), probs, idxs, iou_threshold, self.score_method, self.box_method, True)
        # >= 40000 boxes (nms.py:222-238): per-class clusters on the original coordinates
        kept = torch.zeros_like(scores, dtype=torch.bool)
        parts = []
        for cid in torch.unique(idxs).tolist():
            sel = (idxs == cid).nonzero().view(-1)
            k, b, s, p, l = ops.fusion_nms(boxes[sel].float(), probs[sel], idxs[sel], iou_threshold,
                                           self.score_method, self.box_method, False)
            parts.append((b, s, p, l))
            kept[sel[k]] = True
        b, s, p, l = (torch.cat([q[i] for q in parts], dim=0) for i in range(4))
        order = s.argsort(descending=True)
        return kept.nonzero().view(-1)[order], b[order], s[order], p[order], l[order]


mynms = MyNMS(method=None)

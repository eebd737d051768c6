"""detectron2.modeling.proposal_generator.proposal_utils: add_ground_truth_to_proposals and
find_top_rpn_proposals (restated from detectron2 0.5; the reference passes Instances carrying gt_boxes
to add_ground_truth_to_proposals, clip_roi_heads.py:347-348, which the single-image helper unwraps)."""
import math
from typing import List, Tuple

import torch

from detectron2.layers import batched_nms, cat
from detectron2.structures import Boxes, Instances


def find_top_rpn_proposals(proposals: List[torch.Tensor], pred_objectness_logits: List[torch.Tensor],
                           image_sizes: List[Tuple[int, int]], nms_thresh: float, pre_nms_topk: int,
                           post_nms_topk: int, min_box_size: float, training: bool):
    num_images = len(image_sizes)
    device = proposals[0].device
    topk_scores, topk_proposals, level_ids = [], [], []
    batch_idx = torch.arange(num_images, device=device)
    for level_id, (proposals_i, logits_i) in enumerate(zip(proposals, pred_objectness_logits)):
        Hi_Wi_A = logits_i.shape[1]
        num_proposals_i = min(Hi_Wi_A, pre_nms_topk)
        # the 0.5 source sorts and slices; ties of equal logits are resolved in index order (stable)
        logits_i, idx = logits_i.sort(descending=True, dim=1, stable=True)
        topk_scores_i = logits_i.narrow(1, 0, num_proposals_i)
        topk_idx = idx.narrow(1, 0, num_proposals_i)
        topk_proposals_i = proposals_i[batch_idx[:, None], topk_idx]
        topk_proposals.append(topk_proposals_i)
        topk_scores.append(topk_scores_i)
        level_ids.append(torch.full((num_proposals_i,), level_id, dtype=torch.int64, device=device))
    topk_scores = cat(topk_scores, dim=1)
    topk_proposals = cat(topk_proposals, dim=1)
    level_ids = cat(level_ids, dim=0)
    results = []
    for n, image_size in enumerate(image_sizes):
        boxes = Boxes(topk_proposals[n])
        scores_per_img = topk_scores[n]
        lvl = level_ids
        valid_mask = torch.isfinite(boxes.tensor).all(dim=1) & torch.isfinite(scores_per_img)
        if not valid_mask.all():
            if training:
                raise FloatingPointError("Predicted boxes or scores contain Inf/NaN. Training has diverged.")
            boxes = boxes[valid_mask]
            scores_per_img = scores_per_img[valid_mask]
            lvl = lvl[valid_mask]
        boxes.clip(image_size)
        keep = boxes.nonempty(threshold=min_box_size)
        if keep.sum().item() != len(boxes):
            boxes, scores_per_img, lvl = boxes[keep], scores_per_img[keep], lvl[keep]
        keep = batched_nms(boxes.tensor, scores_per_img, lvl, nms_thresh)
        keep = keep[:post_nms_topk]
        res = Instances(image_size)
        res.proposal_boxes = boxes[keep]
        res.objectness_logits = scores_per_img[keep]
        results.append(res)
    return results


def add_ground_truth_to_proposals(gt, proposals):
    assert gt is not None
    assert len(proposals) == len(gt)
    if len(proposals) == 0:
        return proposals
    return [add_ground_truth_to_proposals_single_image(gt_i, proposals_i) for gt_i, proposals_i in zip(gt, proposals)]


def add_ground_truth_to_proposals_single_image(gt, proposals):
    gt_boxes = gt.gt_boxes if isinstance(gt, Instances) else gt
    device = proposals.objectness_logits.device
    gt_logit_value = math.log((1.0 - 1e-10) / (1 - (1.0 - 1e-10)))
    gt_logits = gt_logit_value * torch.ones(len(gt_boxes), device=device)
    gt_proposal = Instances(proposals.image_size)
    gt_proposal.proposal_boxes = gt_boxes
    gt_proposal.objectness_logits = gt_logits
    return Instances.cat([proposals, gt_proposal])

"""detectron2.modeling.box_regression: Box2BoxTransform, _dense_box_regression_loss (restated from
detectron2 0.5 modeling/box_regression.py; smooth_l1 branch only, the one the reference configures)."""
import math
from typing import List, Tuple

import torch
from fvcore.nn import smooth_l1_loss

from detectron2.layers import cat
from detectron2.structures import Boxes

_DEFAULT_SCALE_CLAMP = math.log(1000.0 / 16)


class Box2BoxTransform(object):
    def __init__(self, weights: Tuple[float, float, float, float], scale_clamp: float = _DEFAULT_SCALE_CLAMP):
        self.weights = weights
        self.scale_clamp = scale_clamp

    def get_deltas(self, src_boxes, target_boxes):
        assert isinstance(src_boxes, torch.Tensor), type(src_boxes)
        assert isinstance(target_boxes, torch.Tensor), type(target_boxes)
        src_widths = src_boxes[:, 2] - src_boxes[:, 0]
        src_heights = src_boxes[:, 3] - src_boxes[:, 1]
        src_ctr_x = src_boxes[:, 0] + 0.5 * src_widths
        src_ctr_y = src_boxes[:, 1] + 0.5 * src_heights
        target_widths = target_boxes[:, 2] - target_boxes[:, 0]
        target_heights = target_boxes[:, 3] - target_boxes[:, 1]
        target_ctr_x = target_boxes[:, 0] + 0.5 * target_widths
        target_ctr_y = target_boxes[:, 1] + 0.5 * target_heights
        wx, wy, ww, wh = self.weights
        dx = wx * (target_ctr_x - src_ctr_x) / src_widths
        dy = wy * (target_ctr_y - src_ctr_y) / src_heights
        dw = ww * torch.log(target_widths / src_widths)
        dh = wh * torch.log(target_heights / src_heights)
        deltas = torch.stack((dx, dy, dw, dh), dim=1)
        assert (src_widths > 0).all().item(), "Input boxes to Box2BoxTransform are not valid!"
        return deltas

    def apply_deltas(self, deltas, boxes):
        deltas = deltas.float()
        boxes = boxes.to(deltas.dtype)
        widths = boxes[:, 2] - boxes[:, 0]
        heights = boxes[:, 3] - boxes[:, 1]
        ctr_x = boxes[:, 0] + 0.5 * widths
        ctr_y = boxes[:, 1] + 0.5 * heights
        wx, wy, ww, wh = self.weights
        dx = deltas[:, 0::4] / wx
        dy = deltas[:, 1::4] / wy
        dw = deltas[:, 2::4] / ww
        dh = deltas[:, 3::4] / wh
        dw = torch.clamp(dw, max=self.scale_clamp)
        dh = torch.clamp(dh, max=self.scale_clamp)
        pred_ctr_x = dx * widths[:, None] + ctr_x[:, None]
        pred_ctr_y = dy * heights[:, None] + ctr_y[:, None]
        pred_w = torch.exp(dw) * widths[:, None]
        pred_h = torch.exp(dh) * heights[:, None]
        x1 = pred_ctr_x - 0.5 * pred_w
        y1 = pred_ctr_y - 0.5 * pred_h
        x2 = pred_ctr_x + 0.5 * pred_w
        y2 = pred_ctr_y + 0.5 * pred_h
        pred_boxes = torch.stack((x1, y1, x2, y2), dim=-1)
        return pred_boxes.reshape(deltas.shape)


def _dense_box_regression_loss(anchors: List[Boxes], box2box_transform: Box2BoxTransform,
                               pred_anchor_deltas: List[torch.Tensor], gt_boxes: List[torch.Tensor],
                               fg_mask: torch.Tensor, box_reg_loss_type="smooth_l1", smooth_l1_beta=0.0):
    anchors = type(anchors[0]).cat(anchors).tensor
    assert box_reg_loss_type == "smooth_l1", "only the configured loss type is restated"
    gt_anchor_deltas = [box2box_transform.get_deltas(anchors, k) for k in gt_boxes]
    gt_anchor_deltas = torch.stack(gt_anchor_deltas)
    return smooth_l1_loss(cat(pred_anchor_deltas, dim=1)[fg_mask], gt_anchor_deltas[fg_mask],
                          beta=smooth_l1_beta, reduction="sum")

"""The two distillation losses of the path's consumers as differentiable single-pass reductions.

    roi_distillation_loss   FastRCNNOutputLayers.losses, "C boxes private" (coin/modeling/roi_heads/fast_rcnn.py:541-545):
                            KLDivLoss(reduction='mean')(log(softmax(scores_c) + 1e-7), gt_probs)
    rpn_distillation_loss   DualTeacherRPN.losses(only_distillation=True) (coin/modeling/proposal_generator/rpn.py:326-340)
                            on the anchors whose distillation label is > 0, teacher_probs from rpn.py:95-98

The reference builds each from ~10 elementwise launches, a boolean-mask gather (host sync) and a reduction; here the
forward is one kernel and the backward one kernel, with the anchor mask and the live C-box count read on the device.
"""
from typing import Optional

import torch

from . import ops


class _RoiDistill(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scores, gt_probs, n_dev):
        ctx.save_for_backward(scores, gt_probs)
        ctx.n_dev = n_dev
        return ops.kl_distill_roi_fwd(scores, gt_probs, n_dev)

    @staticmethod
    def backward(ctx, grad_loss):
        scores, gt_probs = ctx.saved_tensors
        return ops.kl_distill_roi_bwd(scores, gt_probs, grad_loss, ctx.n_dev).to(scores.dtype), None, None


def roi_distillation_loss(scores_c: torch.Tensor, gt_probs: torch.Tensor, weight: float = 1.0,
                          n_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    """scores_c: [n, K+1] class logits of the private (C) boxes; gt_probs: [n, K+1] teacher probabilities.
    n_dev: optional device int32 live row count (rows beyond it contribute nothing and get zero gradient)."""
    return _RoiDistill.apply(scores_c, gt_probs, n_dev) * weight


class _RpnDistill(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, teacher):
        loss, n_valid = ops.kl_distill_rpn_fwd(logits, labels, teacher)
        ctx.save_for_backward(logits, labels, teacher, n_valid)
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        logits, labels, teacher, n_valid = ctx.saved_tensors
        return ops.kl_distill_rpn_bwd(logits, labels, teacher, n_valid, grad_loss).to(logits.dtype), None, None


def rpn_distillation_loss(pred_objectness_logits: torch.Tensor, distillation_labels: torch.Tensor, teacher_probs: torch.Tensor,
                          weight: float = 1.0) -> torch.Tensor:
    """pred_objectness_logits / distillation_labels / teacher_probs: [N, A] (or flat). Zero when no anchor is labelled > 0
    (the reference then leaves the loss out of its dict, rpn.py:336-339)."""
    return _RpnDistill.apply(pred_objectness_logits, distillation_labels, teacher_probs) * weight

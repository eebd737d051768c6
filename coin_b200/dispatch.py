"""Dispatcher-level drop-in: the CUDA kernels of ``torchvision::roi_align``, ``torchvision::_roi_align_backward`` and
``torchvision::nms`` replaced by libcoinops (SURVEY.md 8b "torch dispatcher schemas underneath").

The reference never calls these operators by name: it reaches them through detectron2 0.5
(``detectron2.layers.ROIAlign`` -> ``torchvision.ops.roi_align``; ``detectron2.layers.batched_nms`` ->
``torchvision.ops.boxes.batched_nms`` -> ``torch.ops.torchvision.nms``; clip_roi_heads.py:51-63,142-176,
fast_rcnn.py:164, nms.py:207). After ``coin_b200.patch()`` an UNMODIFIED detectron2 / torchvision caller runs this
library's kernels for CUDA tensors: torchvision's own Autograd and Autocast wrappers stay in place (they re-dispatch to
the CUDA key, which is what is overridden here), so ``roi_align(...).backward()`` lands in ``coin_roi_align_bwd``.

    import coin_b200; coin_b200.patch()        # once per process, after `import torchvision`
    coin_b200.unpatch()                        # restores torchvision's kernels

Only the CUDA dispatch key is touched; CPU tensors keep torchvision's CPU kernels (this library has no CPU path).
"""
import warnings
from typing import Optional

import torch

from . import ops

_handle: Optional["torch.library.Library"] = None


def _tv_roi_align(input, rois, spatial_scale, pooled_height, pooled_width, sampling_ratio, aligned):
    # torchvision/csrc/ops/cuda/roi_align_kernel.cu roi_align_forward_kernel: rois [K,5], output [K,C,PH,PW] of input.dtype
    if rois.dim() != 2 or rois.size(1) != 5:
        raise RuntimeError("rois must have shape as Tensor[K, 5]")
    if input.dtype not in (torch.float32, torch.float16):
        raise RuntimeError(f"coin_b200: roi_align supports fp32 / fp16 maps, got {input.dtype}")
    ph, pw = int(pooled_height), int(pooled_width)
    if rois.size(0) == 0:
        return torch.zeros((0, input.size(1), ph, pw), dtype=input.dtype, device=input.device)
    nhwc = ops.to_nhwc_f32(input)
    rois = rois.to(torch.float32)
    # (launch order + size split: a map-sized RoI must not become a 1-ms CTA of the register-tile kernel)
    return ops.roi_align_forward_planned([nhwc], (float(spatial_scale),), rois, (ph, pw), int(sampling_ratio), bool(aligned),
                                         input.dtype, plan=ops.roi_launch_plan(rois, float(spatial_scale))[1])


def _tv_roi_align_backward(grad, rois, spatial_scale, pooled_height, pooled_width, batch_size, channels, height, width,
                           sampling_ratio, aligned):
    shape = (int(batch_size), int(channels), int(height), int(width))
    if grad.dtype not in (torch.float32, torch.float16):
        raise RuntimeError(f"coin_b200: _roi_align_backward supports fp32 / fp16 gradients, got {grad.dtype}")
    if grad.numel() == 0:
        return torch.zeros(shape, dtype=grad.dtype, device=grad.device)
    return ops.roi_align_backward(grad, [shape], (float(spatial_scale),), rois.to(torch.float32), None,
                                  (int(pooled_height), int(pooled_width)), int(sampling_ratio), bool(aligned),
                                  [grad.dtype])[0]


def _tv_nms(dets, scores, iou_threshold):
    # torchvision/csrc/ops/cuda/nms_kernel.cu: kept indices of `dets`, int64, by descending score
    if dets.dim() != 2 or dets.size(1) != 4:
        raise RuntimeError(f"boxes should be a 2d tensor of shape [N, 4], got {tuple(dets.shape)}")
    if scores.dim() != 1 or scores.size(0) != dets.size(0):
        raise RuntimeError("boxes and scores should have same number of elements in dimension 0")
    if dets.numel() == 0:
        return torch.empty((0,), dtype=torch.int64, device=dets.device)
    return ops.nms(dets.to(torch.float32), scores, float(iou_threshold))


def patch() -> None:
    """Route torchvision's CUDA ``roi_align`` / ``_roi_align_backward`` / ``nms`` to libcoinops (idempotent)."""
    global _handle
    if _handle is not None:
        return
    import torchvision  # noqa: F401  (defines the torchvision:: schemas and its own kernels)
    lib = torch.library.Library("torchvision", "IMPL")
    with warnings.catch_warnings():   # "Overriding a previously registered kernel": that is the point
        warnings.simplefilter("ignore", UserWarning)
        lib.impl("roi_align", _tv_roi_align, "CUDA", allow_override=True)
        lib.impl("_roi_align_backward", _tv_roi_align_backward, "CUDA", allow_override=True)
        lib.impl("nms", _tv_nms, "CUDA", allow_override=True)
    _handle = lib


def unpatch() -> None:
    """Undo ``patch()``: torchvision's own CUDA kernels are dispatched again."""
    global _handle
    if _handle is not None:
        _handle._destroy()
        _handle = None


def is_patched() -> bool:
    return _handle is not None

"""Pascal-VOC AP for one class on the device: the ``voc_eval`` / ``voc_ap`` pair of
``coin/evaluation/cloud_pascal_voc_evaluation.py:173-319`` on tensors instead of text files.

The reference writes every detection to a per-class text file, parses it back, sorts by confidence and walks the
detections one by one in numpy (one IoU vector per detection). Here the confidence order comes from ``coin_argsort_desc``,
the TP / FP marking from ``coin_voc_match`` (one thread per detection, the "already detected" flag resolved by an
atomicMin on the detection rank), and the cumulative sums and the AP integral are a handful of float64 tensor operations
on the device. Policy on exact confidence ties: stable (lower index first); ``order=`` replays another order, e.g. the
``np.argsort(-confidence)`` of a reference run.
"""
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops
from ._lib import check, lib


def pack_ground_truth(gt_boxes_per_image: Sequence[torch.Tensor], gt_difficult_per_image: Sequence[torch.Tensor], device):
    """Per-image lists -> (gt_boxes float64 [ng,4], gt_offsets int32 [n_img+1], gt_difficult uint8 [ng]) on ``device``."""
    counts = [int(b.shape[0]) for b in gt_boxes_per_image]
    offsets = torch.zeros(len(counts) + 1, dtype=torch.int32)
    offsets[1:] = torch.tensor(counts, dtype=torch.int32).cumsum(0) if counts else offsets[1:]
    boxes = (torch.cat([b.reshape(-1, 4).to(torch.float64) for b in gt_boxes_per_image]) if counts
             else torch.zeros((0, 4), dtype=torch.float64))
    diff = (torch.cat([d.reshape(-1).to(torch.uint8) for d in gt_difficult_per_image]) if counts
            else torch.zeros((0,), dtype=torch.uint8))
    return boxes.to(device), offsets.to(device), diff.to(device)


def voc_match(det_image: torch.Tensor, det_boxes: torch.Tensor, order: torch.Tensor, gt_boxes: torch.Tensor,
              gt_offsets: torch.Tensor, gt_difficult: torch.Tensor, ovthresh: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """(tp, fp): float64 [nd] flags in confidence order."""
    dev = det_boxes.device
    nd, ng = int(det_boxes.shape[0]), int(gt_boxes.shape[0])
    di = ops._cuda(det_image, "det_image").to(torch.int32).contiguous()
    db = ops._cuda(det_boxes, "det_boxes").to(torch.float64).contiguous()
    od = ops._i64c(order, "order")
    gb = ops._cuda(gt_boxes, "gt_boxes").to(torch.float64).contiguous()
    go = ops._cuda(gt_offsets, "gt_offsets").to(torch.int32).contiguous()
    gd = ops._cuda(gt_difficult, "gt_difficult").to(torch.uint8).contiguous()
    tp = torch.zeros((nd,), dtype=torch.float64, device=dev)
    fp = torch.zeros((nd,), dtype=torch.float64, device=dev)
    ws = ops._workspace(lib.coin_voc_match_workspace_bytes(nd, ng), dev)
    check(lib.coin_voc_match(ops._ptr(di), ops._ptr(db), ops._ptr(od), nd, ops._ptr(gb), ops._ptr(go), ops._ptr(gd), ng,
                             float(ovthresh), ops._ptr(tp), ops._ptr(fp), ops._ptr(ws), ws.numel(), ops._stream()))
    return tp, fp


def argsort_desc(scores: torch.Tensor) -> torch.Tensor:
    """Stable descending argsort of fp32 scores on the device (ties: lower index first)."""
    s = ops._f32c(scores, "scores").reshape(-1)
    order = torch.empty((s.numel(),), dtype=torch.int64, device=s.device)
    ws = ops._workspace(lib.coin_argsort_desc_workspace_bytes(s.numel()), s.device)
    check(lib.coin_argsort_desc(ops._ptr(s), s.numel(), ops._ptr(order), ops._ptr(ws), ws.numel(), ops._stream()))
    return order


def voc_ap(rec: torch.Tensor, prec: torch.Tensor, use_07_metric: bool = False) -> float:
    """cloud_pascal_voc_evaluation.py:173-202 on device tensors (float64)."""
    if use_07_metric:
        ap = 0.0
        for t in np.arange(0.0, 1.1, 0.1):
            m = rec >= float(t)
            p = float(prec[m].max()) if bool(m.any()) else 0.0
            ap = ap + p / 11.0
        return ap
    z, o = rec.new_zeros(1), rec.new_ones(1)
    mrec = torch.cat((z, rec, o))
    mpre = torch.cat((z, prec, z))
    mpre = torch.flip(torch.cummax(torch.flip(mpre, (0,)), dim=0).values, (0,))     # the precision envelope
    i = torch.nonzero(mrec[1:] != mrec[:-1]).reshape(-1)
    return float(((mrec[i + 1] - mrec[i]) * mpre[i + 1]).sum())


def voc_eval_class(det_image: torch.Tensor, det_conf: torch.Tensor, det_boxes: torch.Tensor, gt_boxes: torch.Tensor,
                   gt_offsets: torch.Tensor, gt_difficult: torch.Tensor, ovthresh: float = 0.5, use_07_metric: bool = False,
                   order: Optional[torch.Tensor] = None):
    """``voc_eval`` for one class: returns (rec, prec, ap) like the reference (rec / prec as float64 device tensors)."""
    if order is None:
        order = argsort_desc(det_conf)
    tp, fp = voc_match(det_image, det_boxes, order, gt_boxes, gt_offsets, gt_difficult, ovthresh)
    tp, fp = torch.cumsum(tp, 0), torch.cumsum(fp, 0)
    # tensor / tensor: an IEEE division per element like numpy's (a Python-scalar divisor becomes a multiplication by
    # its reciprocal on the device, one ulp off)
    npos = (gt_difficult == 0).sum().to(torch.float64).expand_as(tp)
    rec = tp / npos
    prec = tp / torch.clamp(tp + fp, min=float(np.finfo(np.float64).eps))
    return rec, prec, voc_ap(rec, prec, use_07_metric)

"""Functional wrappers: torch CUDA tensors in, torch CUDA tensors out, every call through the C ABI.

Nothing here computes on the host or through torch operators except allocation, dtype casts that
the reference itself performs (``boxes.float()``, ``deltas.float()``) and narrowing a worst-case
output buffer to the device-side count. A CPU tensor is an error (no fallback).
"""
import ctypes
import math
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import check, lib

_SCALE_CLAMP = math.log(1000.0 / 16)


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _cuda(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or t.device.type != "cuda":
        raise RuntimeError(f"coin_b200: `{name}` must be a CUDA tensor (there is no CPU fallback)")
    return t


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    return _cuda(t, name).to(torch.float32).contiguous()


def _i64c(t: torch.Tensor, name: str) -> torch.Tensor:
    return _cuda(t, name).to(torch.int64).contiguous()


def _boxes(t: torch.Tensor, name: str) -> torch.Tensor:
    t = _f32c(t, name)
    if t.dim() != 2 or t.shape[-1] != 4:
        if t.numel() == 0:
            return t.reshape(0, 4)
        raise ValueError(f"coin_b200: `{name}` must have shape [n, 4], got {tuple(t.shape)}")
    return t


def _dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return _lib.F32
    if dt == torch.float16:
        return _lib.F16
    raise TypeError(f"coin_b200: unsupported dtype {dt} (fp32 and fp16 are supported)")


def _event_pair(events):
    if events is None:
        return None
    # inside a CUDA-graph capture the pair becomes two event-record NODES (external events): every replay
    # re-records them, so elapsed_time() after a replay is the kernel's duration inside the graph
    ext = torch.cuda.is_current_stream_capturing()
    pair = (torch.cuda.Event(enable_timing=True, external=ext), torch.cuda.Event(enable_timing=True, external=ext))
    pair[0].record()
    events.append(pair)
    return pair


def _event_close(pair):
    if pair is not None:
        pair[1].record()


def _workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty((max(int(nbytes), 256),), dtype=torch.uint8, device=device)


def _count(t: torch.Tensor) -> torch.Tensor:
    """A device-side count: one int32 element (typically a view into a larger counts tensor)."""
    if t.device.type != "cuda" or t.dtype != torch.int32 or t.numel() != 1:
        raise ValueError("coin_b200: a device count must be a 1-element int32 CUDA tensor")
    return t


# ------------------------------------------------------------------------------------------------
# ROIAlign
# ------------------------------------------------------------------------------------------------
def to_nhwc_f32(x: torch.Tensor) -> torch.Tensor:
    """[N,C,H,W] fp32/fp16 -> fp32 [N,H,W,C] contiguous (the layout the gather kernels read)."""
    _cuda(x, "input")
    n, c, h, w = x.shape
    if x.dtype == torch.float32 and x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous():
        return x.permute(0, 2, 3, 1)  # already channel-last in memory: zero copy
    x = x.contiguous()
    out = torch.empty((n, h, w, c), dtype=torch.float32, device=x.device)
    check(lib.coin_nchw_to_nhwc_f32(_ptr(x), _dtype_code(x.dtype), _ptr(out), n, c, h, w, _stream()))
    return out


def _levels(feats_nhwc: Sequence[torch.Tensor], scales: Sequence[float]):
    arr = (_lib.CoinLevel * len(feats_nhwc))()
    for i, (f, s) in enumerate(zip(feats_nhwc, scales)):
        arr[i].feat_nhwc = f.data_ptr()
        arr[i].H, arr[i].W = int(f.shape[1]), int(f.shape[2])
        arr[i].spatial_scale = float(s)
    return arr


ORDER_MAX_ROIS = 8192          # coin_roi_launch_order's limit
ORDER_MIN_ROIS = 512           # below ~1 wave of CTAs there is no tail to fill


def roi_launch_order(rois: torch.Tensor, k_dev: Optional[torch.Tensor] = None, small_pct: int = 20,
                     big_pct: int = 0) -> Optional[torch.Tensor]:
    """Launch order for roi_align_forward / roi_align_backward over the same ``rois`` (scheduling only): the smallest
    ``small_pct`` % of the RoIs go last, the largest ``big_pct`` % first (default 0: on the bench shape largest-first costs the
    forward 4 %, which wants load-heavy and store-heavy CTAs mixed; it is for launches of few waves that may hold a map-sized
    RoI). Returns None where ordering does not pay (few or very many RoIs)."""
    rois = _f32c(rois, "rois")
    k = rois.shape[0]
    if k < ORDER_MIN_ROIS or k > ORDER_MAX_ROIS or (small_pct <= 0 and big_pct <= 0):
        return None
    perm = torch.empty((k,), dtype=torch.int32, device=rois.device)
    check(lib.coin_roi_launch_order(_ptr(rois), k, _ptr(None if k_dev is None else _count(k_dev)), int(small_pct),
                                    int(big_pct), _ptr(perm), _stream()))
    return perm


def roi_launch_plan(rois: torch.Tensor, scale: float, k_dev: Optional[torch.Tensor] = None, small_pct: int = 20,
                    big_pct: int = 0):
    """Launch order + size split in ONE launch (coin_roi_launch_plan): returns (perm, plan) where perm is the launch order over
    all live RoIs (for the backward) and plan = (perm, perm_divert, counts) is what roi_align_forward_planned takes: the main
    kernel pools perm[:counts[0]], the separable kernel the map-sized RoIs perm_divert[:counts[1]]. None, None beyond 8192 RoIs."""
    rois = _f32c(rois, "rois")
    k = rois.shape[0]
    if k > ORDER_MAX_ROIS or k == 0:
        return None, None
    stride = 1.0 / float(scale)
    perm = torch.empty((k,), dtype=torch.int32, device=rois.device)
    cap = min(BIG_ROI_CAP, k)
    pd = torch.empty((cap,), dtype=torch.int32, device=rois.device)
    counts = torch.empty((2,), dtype=torch.int32, device=rois.device)     # (written by the kernel: no fill launch)
    if k < ORDER_MIN_ROIS:
        small_pct = big_pct = 0              # below ~1 wave of CTAs there is no tail to fill: only the split
    check(lib.coin_roi_launch_plan(_ptr(rois), k, _ptr(None if k_dev is None else _count(k_dev)), int(small_pct), int(big_pct),
                                   BIG_ROI_CELLS * stride * stride, BIG_ROI_SIDE * stride, cap, _ptr(perm), _ptr(pd),
                                   _ptr(counts), _stream()))
    return perm, (perm, pd, counts)


BIG_ROI_CELLS = 1000           # a RoI covering more feature cells than this, or wider / taller than BIG_ROI_SIDE cells, is "big":
BIG_ROI_SIDE = 40              # a register-tile CTA walks it for 0.2 - 1 ms (tools/c_box_probe.py); the separable kernel pools it.
#                                (500 x 500 px at stride 16 - the largest box of the Foggy size law, 47 us - stays below both.)
BIG_ROI_CAP = 32               # ... up to this many per call (the capacity of that second, normally empty, launch)


def roi_split_by_area(rois: torch.Tensor, k_dev: Optional[torch.Tensor], area_thr: float, side_thr: float = float("inf"),
                      order: Optional[torch.Tensor] = None, big_cap: Optional[int] = None):
    """(perm_small, perm_big, counts): index lists of the "big" RoIs (box area > ``area_thr`` px^2 or a side > ``side_thr``
    px) and of the rest (in input order, or in the order of ``order``), with their device lengths counts[0:1] (small),
    counts[1:2] (big) - for pooling the two subsets with different kernels into one output."""
    rois = _f32c(rois, "rois")
    k = rois.shape[0]
    big_cap = k if big_cap is None else min(int(big_cap), k)
    ps = torch.empty((max(k, 1),), dtype=torch.int32, device=rois.device)
    pb = torch.empty((max(big_cap, 1),), dtype=torch.int32, device=rois.device)
    counts = torch.zeros((2,), dtype=torch.int32, device=rois.device)
    check(lib.coin_roi_split_by_area(_ptr(rois), k, _ptr(None if k_dev is None else _count(k_dev)), float(area_thr), min(float(side_thr), 3.0e38),
                                     big_cap, _ptr(order), _ptr(ps), _ptr(pb), _ptr(counts), _stream()))
    return ps, pb, counts


def roi_align_forward_planned(feats_nhwc: Sequence[torch.Tensor], scales: Sequence[float], rois: torch.Tensor,
                              output_size: Tuple[int, int], sampling_ratio: int, aligned: bool, out_dtype: torch.dtype,
                              order: Optional[torch.Tensor] = None, k_dev: Optional[torch.Tensor] = None,
                              events: Optional[list] = None, plan=None, big_stream=None, return_plan: bool = False):
    """Single-level roi_align_forward that is robust to map-sized RoIs: the RoIs are split on the device by size
    (BIG_ROI_CELLS / BIG_ROI_SIDE feature cells); the big ones are pooled by the separable kernel, the rest by the default
    (register-tile) kernel in the launch order ``order``; one output. plan: a precomputed roi_split_by_area result.
    big_stream: optional side stream for the (normally empty) launch of the big RoIs, so that it does not sit in front of the
    main launch; the current stream waits for it before returning."""
    if _lib.get_option("COIN_ROI_EXACT", 0) != 0:     # the parity kernels pool every RoI the same way
        out = roi_align_forward(feats_nhwc, scales, rois, None, output_size, sampling_ratio, aligned, out_dtype, events, k_dev)
        return (out, None) if return_plan else out
    stride = 1.0 / float(scales[0])
    if plan is None:
        plan = roi_split_by_area(rois, k_dev, BIG_ROI_CELLS * stride * stride, BIG_ROI_SIDE * stride, order, BIG_ROI_CAP)
    p_small, p_big, cnt = plan
    k, c = rois.shape[0], int(feats_nhwc[0].shape[3])
    out = torch.empty((k, c) + tuple(output_size), dtype=out_dtype, device=rois.device)
    def pool_big():     # the big list: a launch sized for its capacity (a few CTAs), normally empty
        with _lib.options(COIN_ROI_REG=0):
            roi_align_forward(feats_nhwc, scales, rois, None, output_size, sampling_ratio, aligned, out_dtype, k_dev=cnt[1:2],
                              perm=p_big, out=out, k_launch=int(p_big.shape[0]))
    big_done = None
    if big_stream is None:
        pool_big()
    else:
        big_stream.wait_event(torch.cuda.current_stream().record_event())
        with torch.cuda.stream(big_stream):
            pool_big()
            big_done = big_stream.record_event()
    roi_align_forward(feats_nhwc, scales, rois, None, output_size, sampling_ratio, aligned, out_dtype, events=events,
                      k_dev=cnt[0:1], perm=p_small, out=out)
    if big_done is not None:
        torch.cuda.current_stream().wait_event(big_done)
    return (out, plan) if return_plan else out


def roi_align_forward(feats_nhwc: Sequence[torch.Tensor], scales: Sequence[float], rois: torch.Tensor,
                      roi_level: Optional[torch.Tensor], output_size: Tuple[int, int], sampling_ratio: int,
                      aligned: bool, out_dtype: torch.dtype, events: Optional[list] = None,
                      k_dev: Optional[torch.Tensor] = None, perm: Optional[torch.Tensor] = None,
                      out: Optional[torch.Tensor] = None, k_launch: Optional[int] = None) -> torch.Tensor:
    """events: optional list; a (start, end) CUDA-event pair bracketing the kernel launch is appended.
    k_dev: optional device int32 live RoI count (<= K): rows of the result beyond it are not written.
    perm: optional launch order from ``roi_launch_order`` (same rois).
    out: optional preallocated contiguous [K, C, PH, PW] result (e.g. a row range of a larger buffer).
    k_launch: with perm, the number of RoI slots to launch (<= K; the live count *k_dev must not exceed it)."""
    rois = _f32c(rois, "rois")
    if rois.dim() != 2 or rois.shape[1] != 5:
        raise ValueError(f"coin_b200: rois must have shape [K, 5], got {tuple(rois.shape)}")
    k, c = rois.shape[0], int(feats_nhwc[0].shape[3])
    ph, pw = output_size
    if out is None:
        out = torch.empty((k, c, ph, pw), dtype=out_dtype, device=rois.device)
    elif tuple(out.shape) != (k, c, ph, pw) or out.dtype != out_dtype or not out.is_contiguous():
        raise ValueError("coin_b200: `out` must be a contiguous [K, C, PH, PW] tensor of the output dtype")
    if roi_level is not None:
        roi_level = _cuda(roi_level, "roi_level").to(torch.int32).contiguous()
    pair = _event_pair(events)
    if perm is not None:
        check(lib.coin_roi_align_fwd_ord(_levels(feats_nhwc, scales), len(feats_nhwc), _ptr(rois), _ptr(roi_level),
                                         _ptr(out), _dtype_code(out_dtype), c, k if k_launch is None else min(int(k_launch), k),
                                         ph, pw, int(sampling_ratio),
                                         int(bool(aligned)), _ptr(None if k_dev is None else _count(k_dev)),
                                         _ptr(perm), _stream()))
    elif k_dev is None:
        check(lib.coin_roi_align_fwd(_levels(feats_nhwc, scales), len(feats_nhwc), _ptr(rois), _ptr(roi_level),
                                     _ptr(out), _dtype_code(out_dtype), c, k, ph, pw, int(sampling_ratio),
                                     int(bool(aligned)), _stream()))
    else:
        check(lib.coin_roi_align_fwd_dev(_levels(feats_nhwc, scales), len(feats_nhwc), _ptr(rois), _ptr(roi_level),
                                         _ptr(out), _dtype_code(out_dtype), c, k, ph, pw, int(sampling_ratio),
                                         int(bool(aligned)), _ptr(_count(k_dev)), _stream()))
    _event_close(pair)
    return out


def roi_align_backward(grad_out: torch.Tensor, shapes: Sequence[Tuple[int, int, int, int]], scales: Sequence[float],
                       rois: torch.Tensor, roi_level: Optional[torch.Tensor], output_size: Tuple[int, int],
                       sampling_ratio: int, aligned: bool, out_dtypes: Sequence[torch.dtype],
                       events: Optional[list] = None, perm: Optional[torch.Tensor] = None, plan=None,
                       big_stream=None, zeroed: Optional[Sequence[torch.Tensor]] = None) -> List[torch.Tensor]:
    """Returns one NCHW gradient per level (shape ``shapes[i]`` = (N,C,H,W), dtype ``out_dtypes[i]``).
    perm: launch order (roi_launch_order). plan: a roi_split_by_area result of the forward (single level): the map-sized
    RoIs' gradients are scattered by the separable kernel (on ``big_stream`` if given), the rest by the default kernel.
    zeroed: one ZERO-FILLED fp32 [N,H,W,C] accumulation buffer per level, ready on the current stream (a caller that zeroes
    them ahead of time, beside other work, takes the fill off the path between forward and backward); consumed."""
    grad_out = _cuda(grad_out, "grad_out").contiguous()
    rois = _f32c(rois, "rois")
    k, c = rois.shape[0], shapes[0][1]
    ph, pw = output_size
    if zeroed is not None:
        bufs = list(zeroed)
        for buf, (n, cc, h, w) in zip(bufs, shapes):
            if tuple(buf.shape) != (n, h, w, cc) or buf.dtype != torch.float32 or not buf.is_contiguous():
                raise ValueError("coin_b200: zeroed buffers must be contiguous fp32 [N,H,W,C] per level")
        if len(bufs) != len(shapes):
            raise ValueError("coin_b200: one zeroed buffer per level")
    else:
        bufs = [torch.zeros((n, h, w, cc), dtype=torch.float32, device=grad_out.device) for (n, cc, h, w) in shapes]
    if roi_level is not None:
        roi_level = _cuda(roi_level, "roi_level").to(torch.int32).contiguous()
    def launch(k_launch, k_dev, order):
        check(lib.coin_roi_align_bwd_ord(_levels(bufs, scales), len(bufs), _ptr(rois), _ptr(roi_level), _ptr(grad_out),
                                         _dtype_code(grad_out.dtype), c, k_launch, ph, pw, int(sampling_ratio),
                                         int(bool(aligned)), _ptr(k_dev), _ptr(order), _stream()))
    big_done = None
    if plan is not None and _lib.get_option("COIN_ROI_EXACT", 0) == 0:
        p_small, p_big, cnt = plan

        def scatter_big():
            with _lib.options(COIN_ROI_REG=0):
                launch(int(p_big.shape[0]), cnt[1:2], p_big)
        if big_stream is None:
            scatter_big()
        else:
            big_stream.wait_event(torch.cuda.current_stream().record_event())      # (the zero-filled buffers exist)
            with torch.cuda.stream(big_stream):
                scatter_big()
                big_done = big_stream.record_event()
        pair = _event_pair(events)
        launch(k, cnt[0:1], p_small)
    else:
        pair = _event_pair(events)
        launch(k, None, perm)
    _event_close(pair)
    if big_done is not None:
        torch.cuda.current_stream().wait_event(big_done)
    outs = []
    for buf, (n, cc, h, w), dt in zip(bufs, shapes, out_dtypes):
        g = torch.empty((n, cc, h, w), dtype=dt, device=buf.device)
        check(lib.coin_nhwc_f32_to_nchw(_ptr(buf), _ptr(g), _dtype_code(dt), n, cc, h, w, _stream()))
        outs.append(g)
    return outs


def roi_pooler_levels(boxes: torch.Tensor, min_level: int, max_level: int, canonical_box_size: int = 224,
                      canonical_level: int = 4) -> torch.Tensor:
    boxes = _boxes(boxes, "boxes")
    out = torch.empty((boxes.shape[0],), dtype=torch.int32, device=boxes.device)
    check(lib.coin_roi_pooler_levels(_ptr(boxes), boxes.shape[0], min_level, max_level, canonical_box_size,
                                     canonical_level, _ptr(out), _stream()))
    return out


# ------------------------------------------------------------------------------------------------
# box codec
# ------------------------------------------------------------------------------------------------
def apply_deltas(deltas: torch.Tensor, boxes: torch.Tensor, weights, scale_clamp: float = _SCALE_CLAMP,
                 clip_to: Optional[Tuple[float, float]] = None) -> torch.Tensor:
    deltas = _f32c(deltas, "deltas")
    boxes = _boxes(boxes, "boxes")
    r = boxes.shape[0]
    if deltas.dim() != 2 or deltas.shape[0] != r or deltas.shape[1] % 4 != 0:
        raise ValueError(f"coin_b200: deltas must have shape [R, 4*k] with R={r}, got {tuple(deltas.shape)}")
    out = torch.empty_like(deltas)
    wx, wy, ww, wh = (float(v) for v in weights)
    h, w = clip_to if clip_to is not None else (0.0, 0.0)
    check(lib.coin_apply_deltas(_ptr(deltas), _ptr(boxes), _ptr(out), r, deltas.shape[1] // 4, wx, wy, ww, wh,
                                float(scale_clamp), int(clip_to is not None), float(h), float(w), _stream()))
    return out


def get_deltas(src: torch.Tensor, tgt: torch.Tensor, weights, check_valid: bool = True) -> torch.Tensor:
    src, tgt = _boxes(src, "src_boxes"), _boxes(tgt, "target_boxes")
    if src.shape != tgt.shape:
        raise ValueError("coin_b200: src_boxes and target_boxes must have the same shape")
    out = torch.empty_like(src)
    flag = torch.zeros((1,), dtype=torch.int32, device=src.device) if check_valid else None
    wx, wy, ww, wh = (float(v) for v in weights)
    check(lib.coin_get_deltas(_ptr(src), _ptr(tgt), _ptr(out), src.shape[0], wx, wy, ww, wh, _ptr(flag), _stream()))
    if check_valid:
        assert int(flag.item()) == 0, "Input boxes to Box2BoxTransform are not valid!"
    return out


def boxes_clip_(boxes: torch.Tensor, image_size: Tuple[float, float]) -> torch.Tensor:
    _cuda(boxes, "boxes")
    if boxes.dtype != torch.float32 or not boxes.is_contiguous():
        raise ValueError("coin_b200: boxes_clip_ needs a contiguous fp32 tensor (it works in place)")
    h, w = image_size
    check(lib.coin_boxes_clip(_ptr(boxes), boxes.numel() // 4, float(h), float(w), _stream()))
    return boxes


def boxes_scale_flip(boxes: torch.Tensor, sx: float, sy: float, flip: str = "no",
                     net_size: Tuple[float, float] = (0.0, 0.0)) -> torch.Tensor:
    boxes = _boxes(boxes, "boxes")
    code = {"no": 0, "horizontal": 1, "vertical": 2}.get(flip)
    if code is None:
        raise NotImplementedError(flip)
    out = torch.empty_like(boxes)
    net_h, net_w = net_size
    check(lib.coin_boxes_scale_flip(_ptr(boxes), _ptr(out), boxes.shape[0], float(sx), float(sy), code, float(net_w),
                                    float(net_h), _stream()))
    return out


def boxes_cxcywh_to_xyxy(boxes: torch.Tensor, image_size: Tuple[float, float], clip: bool = False) -> torch.Tensor:
    """GDINO.resize_boxes (gdino.py:144-160): normalised cxcywh -> pixel xyxy (optionally followed by Boxes.clip)."""
    boxes = _boxes(boxes, "boxes")
    out = torch.empty_like(boxes)
    h, w = image_size
    check(lib.coin_boxes_cxcywh_to_xyxy(_ptr(boxes), _ptr(out), boxes.shape[0], float(h), float(w), int(bool(clip)),
                                        _stream()))
    return out


# ------------------------------------------------------------------------------------------------
# IoU / Matcher
# ------------------------------------------------------------------------------------------------
def pairwise_iou(b1: torch.Tensor, b2: torch.Tensor) -> torch.Tensor:
    b1, b2 = _boxes(b1, "boxes1"), _boxes(b2, "boxes2")
    out = torch.empty((b1.shape[0], b2.shape[0]), dtype=torch.float32, device=b1.device)
    check(lib.coin_pairwise_iou(_ptr(b1), b1.shape[0], _ptr(b2), b2.shape[0], _ptr(out), _stream()))
    return out


def _matcher_cfg(thresholds: Sequence[float], labels: Sequence[int]):
    thr = (ctypes.c_float * len(thresholds))(*[float(t) for t in thresholds])
    lab = (ctypes.c_int8 * len(labels))(*[int(l) for l in labels])
    return thr, lab


def matcher(quality: torch.Tensor, thresholds: Sequence[float], labels: Sequence[int], allow_low_quality: bool,
            return_vals: bool = False):
    q = _f32c(quality, "match_quality_matrix")
    if q.dim() != 2:
        raise ValueError("coin_b200: match_quality_matrix must be 2-D")
    n, m = q.shape
    matches = torch.empty((m,), dtype=torch.int64, device=q.device)
    mlabels = torch.empty((m,), dtype=torch.int8, device=q.device)
    vals = torch.empty((m,), dtype=torch.float32, device=q.device) if return_vals else None
    ws = torch.empty((max(n, 1),), dtype=torch.float32, device=q.device) if allow_low_quality else None
    thr, lab = _matcher_cfg(thresholds, labels)
    check(lib.coin_matcher(_ptr(q), n, m, thr, len(thresholds), lab, int(bool(allow_low_quality)), _ptr(matches),
                           _ptr(mlabels), _ptr(vals), _ptr(ws), _stream()))
    return (matches, mlabels, vals) if return_vals else (matches, mlabels)


def iou_match(gt: torch.Tensor, boxes: torch.Tensor, thresholds: Sequence[float], labels: Sequence[int],
              allow_low_quality: bool, return_vals: bool = False):
    """pairwise_iou(gt, boxes) followed by Matcher, without materialising the matrix."""
    gt, boxes = _boxes(gt, "gt_boxes"), _boxes(boxes, "boxes")
    n, m = gt.shape[0], boxes.shape[0]
    matches = torch.empty((m,), dtype=torch.int64, device=boxes.device)
    mlabels = torch.empty((m,), dtype=torch.int8, device=boxes.device)
    vals = torch.empty((m,), dtype=torch.float32, device=boxes.device) if return_vals else None
    ws = (torch.empty((lib.coin_iou_match_workspace_floats(n, m),), dtype=torch.float32, device=boxes.device)
          if allow_low_quality else None)
    thr, lab = _matcher_cfg(thresholds, labels)
    check(lib.coin_iou_match(_ptr(gt), n, _ptr(boxes), m, thr, len(thresholds), lab, int(bool(allow_low_quality)),
                             _ptr(matches), _ptr(mlabels), _ptr(vals), _ptr(ws), _stream()))
    return (matches, mlabels, vals) if return_vals else (matches, mlabels)


def relabel_roi_(matches: torch.Tensor, labels: torch.Tensor, c_begin: int, c_end: int) -> torch.Tensor:
    check(lib.coin_relabel_roi(_ptr(_cuda(matches, "matches")), _ptr(_cuda(labels, "labels")), matches.numel(),
                               int(c_begin), int(c_end), _stream()))
    return labels


def relabel_rpn_(matches: torch.Tensor, labels: torch.Tensor, len_a: int, len_c: int):
    didx = torch.empty_like(matches)
    dlab = torch.empty_like(labels)
    check(lib.coin_relabel_rpn(_ptr(_cuda(matches, "matches")), _ptr(_cuda(labels, "labels")), matches.numel(),
                               int(len_a), int(len_c), _ptr(didx), _ptr(dlab), _stream()))
    return labels, matches, didx, dlab


def iou_pairs_ge(b1: torch.Tensor, b2: torch.Tensor, thr: float) -> torch.Tensor:
    """== (pairwise_iou(b1, b2) >= thr).nonzero(); int64 [P, 2], row-major order."""
    b1, b2 = _boxes(b1, "boxes1"), _boxes(b2, "boxes2")
    n, m = b1.shape[0], b2.shape[0]
    cap = n * m
    pairs = torch.empty((max(cap, 1), 2), dtype=torch.int64, device=b1.device)
    count = torch.zeros((1,), dtype=torch.int32, device=b1.device)
    nbytes = lib.coin_iou_pairs_workspace_bytes(n, m)
    ws = _workspace(nbytes, b1.device)
    check(lib.coin_iou_pairs_ge(_ptr(b1), n, _ptr(b2), m, float(thr), _ptr(pairs), _ptr(count), cap, _ptr(ws),
                                ws.numel(), _stream()))
    return pairs[: int(count.item())]


# ------------------------------------------------------------------------------------------------
# NMS family
# ------------------------------------------------------------------------------------------------
_STRATEGY = {"plain": _lib.NMS_PLAIN, "trick": _lib.NMS_TRICK, "vanilla": _lib.NMS_VANILLA, "auto": _lib.NMS_AUTO}


def batched_nms(boxes: torch.Tensor, scores: torch.Tensor, idxs: Optional[torch.Tensor], iou_threshold: float,
                strategy: str = "auto", max_keep: int = -1, sync: bool = True):
    """sync=False returns (keep buffer of capacity n, device int32 count) without a host round trip."""
    boxes = _boxes(boxes, "boxes")
    scores = _f32c(scores, "scores")
    n = boxes.shape[0]
    if scores.numel() != n:
        raise ValueError("coin_b200: boxes and scores disagree in length")
    if idxs is not None:
        idxs = _i64c(idxs, "idxs")
    keep = torch.empty((max(n, 1),), dtype=torch.int64, device=boxes.device)
    nkeep = torch.zeros((1,), dtype=torch.int32, device=boxes.device)
    ws = _workspace(lib.coin_nms_workspace_bytes(n), boxes.device)
    check(lib.coin_batched_nms(_ptr(boxes), _ptr(scores), _ptr(idxs), n, float(iou_threshold), _STRATEGY[strategy],
                               int(max_keep), _ptr(keep), _ptr(nkeep), _ptr(ws), ws.numel(), _stream()))
    if not sync:
        return keep, nkeep
    return keep[: int(nkeep.item())]


def rpn_proposals(anchors: Optional[torch.Tensor], deltas: torch.Tensor, logits: torch.Tensor, image_size: Tuple[float, float],
                  pre_nms_topk: int, post_nms_topk: int, nms_thresh: float, min_box_size: float = 0.0,
                  weights=(1.0, 1.0, 1.0, 1.0), scale_clamp: float = _SCALE_CLAMP, sync: bool = True, grid=None):
    """One image, one level of d2 RPN.predict_proposals (decode + find_top_rpn_proposals) in one launch chain.
    anchors: [A,4] - or None with grid = (cell_anchors [ncell,4] (CPU tensor or list), Hf, Wf, stride, offset): the anchors of
    DefaultAnchorGenerator are then generated inside the decode instead of being read.
    Returns (boxes[n,4], logits[n], status) - or, with sync=False, the capacity buffers, the device count and the
    device status word (bit 0: a selected row was non-finite)."""
    deltas = _boxes(deltas, "deltas")
    logits = _f32c(logits, "logits").reshape(-1)
    dev = deltas.device
    if grid is not None:
        cell, hf, wf, stride, offset = grid
        cell = torch.as_tensor(cell, dtype=torch.float32).cpu().reshape(-1, 4).contiguous()
        a = int(hf) * int(wf) * cell.shape[0]
    else:
        anchors = _boxes(anchors, "anchors")
        a = anchors.shape[0]
    if deltas.shape[0] != a or logits.numel() != a:
        raise ValueError("coin_b200: anchors, deltas and logits disagree in length")
    cap = max(min(a, int(pre_nms_topk), int(post_nms_topk)), 1)
    out_boxes = torch.empty((cap, 4), dtype=torch.float32, device=dev)
    out_logits = torch.empty((cap,), dtype=torch.float32, device=dev)
    count = torch.zeros((2,), dtype=torch.int32, device=dev)   # [live rows, status]
    ws = _workspace(lib.coin_rpn_proposals_workspace_bytes(a, int(pre_nms_topk)), dev)
    h, w = image_size
    tail = (int(pre_nms_topk), int(post_nms_topk), float(nms_thresh), float(min_box_size), float(h), float(w), float(weights[0]),
            float(weights[1]), float(weights[2]), float(weights[3]), float(scale_clamp), _ptr(out_boxes), _ptr(out_logits),
            _ptr(count), ctypes.c_void_p(count.data_ptr() + 4), _ptr(ws), ws.numel(), _stream())
    if grid is not None:
        carr = (ctypes.c_float * cell.numel())(*cell.flatten().tolist())
        check(lib.coin_rpn_proposals_grid(carr, cell.shape[0], int(hf), int(wf), float(stride), float(offset), _ptr(deltas),
                                          _ptr(logits), *tail))
    else:
        check(lib.coin_rpn_proposals(_ptr(anchors), _ptr(deltas), _ptr(logits), a, *tail))
    if not sync:
        return out_boxes, out_logits, count
    n, status = count.tolist()
    return out_boxes[:n], out_logits[:n], status


def nms(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float, max_keep: int = -1) -> torch.Tensor:
    return batched_nms(boxes, scores, None, iou_threshold, "plain", max_keep)


_SCORE = {"probEn": _lib.SCORE_PROBEN, "avg": _lib.SCORE_AVG, "max": _lib.SCORE_MAX}
_BOX = {"s-avg": _lib.BOX_SAVG, "avg": _lib.BOX_AVG, "max": _lib.BOX_MAX}


def fusion_nms(boxes: torch.Tensor, probs: torch.Tensor, labels: torch.Tensor, iou_threshold: float,
               score_method: str, box_method: str, per_class_offset: bool = True):
    boxes = _boxes(boxes, "boxes")
    probs = _f32c(probs, "probs")
    labels = _i64c(labels, "idxs")
    n, k1 = boxes.shape[0], int(probs.shape[1])
    dev = boxes.device
    keep = torch.empty((max(n, 1),), dtype=torch.int64, device=dev)
    o_box = torch.empty((max(n, 1), 4), dtype=torch.float32, device=dev)
    o_score = torch.empty((max(n, 1),), dtype=torch.float32, device=dev)
    o_prob = torch.empty((max(n, 1), k1), dtype=torch.float32, device=dev)
    o_cls = torch.empty((max(n, 1),), dtype=torch.int64, device=dev)
    meta = torch.zeros((2,), dtype=torch.int32, device=dev)
    ws = _workspace(lib.coin_fusion_nms_workspace_bytes(n, k1), dev)
    check(lib.coin_fusion_nms(_ptr(boxes), _ptr(probs), _ptr(labels), n, k1, float(iou_threshold), _SCORE[score_method],
                              _BOX[box_method], int(bool(per_class_offset)), _ptr(keep), _ptr(o_box), _ptr(o_score),
                              _ptr(o_prob), _ptr(o_cls), ctypes.c_void_p(meta.data_ptr()),
                              ctypes.c_void_p(meta.data_ptr() + 4), _ptr(ws), ws.numel(), _stream()))
    nk, status = (int(v) for v in meta.tolist())
    if status & 1:
        raise AssertionError("fusion_nms: a cluster mixes classes (nms.py:149,157 assert len(final_class)==1)")
    if status & 2:
        raise AssertionError("fusion_nms: argmax(prob) != label inside a probEn cluster (nms.py:40)")
    return keep[:nk], o_box[:nk], o_score[:nk], o_prob[:nk], o_cls[:nk]


def det_postprocess(boxes: torch.Tensor, scores: torch.Tensor, image_shape: Tuple[int, int], score_thresh: float,
                    nms_thresh: float, topk: int, sync: bool = True):
    """fast_rcnn_inference_single_image. Returns (boxes, scores, probs, classes, roi_index[, count])."""
    boxes = _f32c(boxes, "boxes")
    scores = _f32c(scores, "scores")
    r, k1 = scores.shape
    kreg = boxes.shape[1] // 4
    dev = boxes.device
    cap = int(topk) if topk >= 0 else r * (k1 - 1)
    cap = max(min(cap, r * (k1 - 1)), 1)
    o_box = torch.empty((cap, 4), dtype=torch.float32, device=dev)
    o_score = torch.empty((cap,), dtype=torch.float32, device=dev)
    o_prob = torch.empty((cap, k1), dtype=torch.float32, device=dev)
    o_cls = torch.empty((cap,), dtype=torch.int64, device=dev)
    o_roi = torch.empty((cap,), dtype=torch.int64, device=dev)
    count = torch.zeros((1,), dtype=torch.int32, device=dev)
    ws = _workspace(lib.coin_det_postprocess_workspace_bytes(r, k1), dev)
    h, w = image_shape
    check(lib.coin_det_postprocess(_ptr(boxes), _ptr(scores), r, k1, kreg, float(h), float(w), float(score_thresh),
                                   float(nms_thresh), int(topk), cap, _ptr(o_box), _ptr(o_score), _ptr(o_prob),
                                   _ptr(o_cls), _ptr(o_roi), _ptr(count), _ptr(ws), ws.numel(), _stream()))
    if not sync:
        return o_box, o_score, o_prob, o_cls, o_roi, count
    n = int(count.item())
    return o_box[:n], o_score[:n], o_prob[:n], o_cls[:n], o_roi[:n]


def match_abc(on_boxes, on_classes, on_scores, off_boxes, off_classes, off_scores, tag: str, iou_thr: float,
              weight_for_box_a: float, sync: bool = True):
    """Index form of match_dual_teacher. Returns dict with a_on, a_off, a_boxes, b_*, c_on, c_off.
    sync=False returns the un-narrowed buffers plus the device int32 ``counts`` = [nA, nB, nC, status]
    (narrow later with ``match_abc_narrow`` once the counts have been read back)."""
    on_boxes, off_boxes = _boxes(on_boxes, "online boxes"), _boxes(off_boxes, "offline boxes")
    on_classes, off_classes = _i64c(on_classes, "online classes"), _i64c(off_classes, "offline classes")
    on_scores, off_scores = _f32c(on_scores, "online scores"), _f32c(off_scores, "offline scores")
    nc, nd = on_boxes.shape[0], off_boxes.shape[0]
    dev = on_boxes.device
    cap = nc * nd + nc + nd
    i32 = lambda n: torch.empty((max(n, 1),), dtype=torch.int32, device=dev)
    a_on, a_off, b_on, b_off = i32(cap), i32(cap), i32(cap), i32(cap)
    c_on, c_off = i32(nc + nd), i32(nc + nd)
    a_box = torch.empty((max(cap, 1), 4), dtype=torch.float32, device=dev)
    b_box = torch.empty((max(cap, 1), 4), dtype=torch.float32, device=dev)
    counts = torch.zeros((8,), dtype=torch.int32, device=dev)
    ws = _workspace(lib.coin_match_abc_workspace_bytes(nc, nd), dev)
    code = {"RCNN": _lib.TAG_RCNN, "RPN": _lib.TAG_RPN}[tag]
    check(lib.coin_match_abc(_ptr(on_boxes), _ptr(on_classes), _ptr(on_scores), nc, _ptr(off_boxes), _ptr(off_classes),
                             _ptr(off_scores), nd, code, float(iou_thr), float(weight_for_box_a), cap, _ptr(a_on),
                             _ptr(a_off), _ptr(a_box), _ptr(b_on), _ptr(b_off), _ptr(b_box), _ptr(c_on), _ptr(c_off),
                             _ptr(counts), _ptr(ws), ws.numel(), _stream()))
    raw = {"a_on": a_on, "a_off": a_off, "a_boxes": a_box, "b_on": b_on, "b_off": b_off, "b_boxes": b_box,
           "c_on": c_on, "c_off": c_off, "counts": counts}
    if not sync:
        return raw
    return match_abc_narrow(raw, counts.tolist())


def match_abc_narrow(raw, counts):
    na, nb, ncc, status, nc_off = (int(v) for v in counts[:5])
    a_on, a_off, a_box, b_on, b_off, b_box, c_on, c_off = (raw[k] for k in (
        "a_on", "a_off", "a_boxes", "b_on", "b_off", "b_boxes", "c_on", "c_off"))
    if status & 16:
        raise AssertionError("match_abc: a cloud self-cluster has a single class (util.py:488 assert)")
    if status & 8:
        raise AssertionError("match_abc: a duplicate group holds several boxes of the matched class "
                             "(trainer.py:382 would desynchronise the common lists)")
    if status & 4:
        raise RuntimeError("match_abc: pair capacity exceeded")
    return {"a_on": a_on[:na].long(), "a_off": a_off[:na].long(), "a_boxes": a_box[:na],
            "b_on": b_on[:nb].long(), "b_off": b_off[:nb].long(), "b_boxes": b_box[:nb],
            "c_on": c_on[nc_off:ncc].long(), "c_off": c_off[:nc_off].long()}


# ------------------------------------------------------------------------------------------------
# sync-free (device-count) variants: worst-case buffers + device int32 lengths, no host round trip
# ------------------------------------------------------------------------------------------------
def concat_rows(segments, width_out: Optional[int] = None, out_cap: Optional[int] = None):
    """segments: list of (tensor [n, w] fp32, count_dev or None, prefix float). Returns (out [cap, width_out],
    device int32 count). == torch.cat of the live prefixes (optionally with a leading prefix column)."""
    n = len(segments)
    arr = (_lib.CoinSeg * n)()
    keep = []
    width_in = int(segments[0][0].shape[1])
    worst = 0
    for i, (t, cnt, prefix) in enumerate(segments):
        t = _f32c(t, "segment")
        keep.append(t)
        arr[i].ptr = t.data_ptr()
        arr[i].count_dev = 0 if cnt is None else _count(cnt).data_ptr()
        arr[i].count = int(t.shape[0])
        arr[i].prefix = float(prefix)
        worst += int(t.shape[0])
    width_out = width_in if width_out is None else int(width_out)
    cap = worst if out_cap is None else int(out_cap)
    dev = keep[0].device
    out = torch.empty((max(cap, 1), width_out), dtype=torch.float32, device=dev)
    count = torch.empty((1,), dtype=torch.int32, device=dev)       # (written by the kernel: no fill launch)
    check(lib.coin_concat_rows(arr, n, width_in, width_out, _ptr(out), cap, _ptr(count), _stream()))
    return out, count


def iou_match_dev(gt: torch.Tensor, n_dev: Optional[torch.Tensor], boxes: torch.Tensor, m_dev: Optional[torch.Tensor],
                  thresholds: Sequence[float], labels: Sequence[int], allow_low_quality: bool):
    """iou_match with device-side live counts of gt rows / columns. Outputs have the capacity of `boxes`."""
    gt, boxes = _boxes(gt, "gt_boxes"), _boxes(boxes, "boxes")
    n, m = gt.shape[0], boxes.shape[0]
    matches = torch.empty((m,), dtype=torch.int64, device=boxes.device)
    mlabels = torch.empty((m,), dtype=torch.int8, device=boxes.device)
    ws = (torch.empty((lib.coin_iou_match_workspace_floats(n, m),), dtype=torch.float32, device=boxes.device)
          if allow_low_quality else None)
    thr, lab = _matcher_cfg(thresholds, labels)
    check(lib.coin_iou_match_dev(_ptr(gt), n, _ptr(None if n_dev is None else _count(n_dev)), _ptr(boxes), m,
                                 _ptr(None if m_dev is None else _count(m_dev)), thr, len(thresholds), lab,
                                 int(bool(allow_low_quality)), _ptr(matches), _ptr(mlabels), _ptr(None), _ptr(ws),
                                 _stream()))
    return matches, mlabels


def relabel_roi_dev_(matches, labels, m_dev, len_a, len_b, len_c):
    check(lib.coin_relabel_roi_dev(_ptr(matches), _ptr(labels), matches.numel(),
                                   _ptr(None if m_dev is None else _count(m_dev)), _ptr(_count(len_a)),
                                   _ptr(_count(len_b)), _ptr(_count(len_c)), _stream()))
    return labels


def relabel_rpn_dev_(matches, labels, len_a, len_c):
    didx = torch.empty_like(matches)
    dlab = torch.empty_like(labels)
    check(lib.coin_relabel_rpn_dev(_ptr(matches), _ptr(labels), matches.numel(), _ptr(_count(len_a)),
                                   _ptr(_count(len_c)), _ptr(didx), _ptr(dlab), _stream()))
    return labels, matches, didx, dlab


def match_abc_fields_dev(on: dict, off: dict, nd_dev: torch.Tensor, tag: str, iou_thr: float, weight_for_box_a: float):
    """Knowledge separation + field gathers for ONE tag with the CLIP-detector detection count on the device.
    Returns (A, B or None, C, counts); see match_abc_fields_both_dev."""
    return match_abc_fields_both_dev(on, off, nd_dev, iou_thr, weight_for_box_a, tags=(tag,))[tag]


def match_abc_fields_both_dev(on: dict, off: dict, nd_dev: torch.Tensor, iou_thr: float, weight_for_box_a: float,
                              tags=("RCNN", "RPN")):
    """Knowledge separation (one launch for all requested tags) + field gathers (one launch per tag), with the
    CLIP-detector detection count on the device.
    on / off: dicts with gt_boxes [n,4], gt_classes int64 [n], scores [n], probs [n,k1] (off is padded to
    its capacity; nd_dev holds the live count). Returns {tag: (A, B or None, C, counts)}: padded field dicts with
    the reference's names and the device int32 counts [nA, nB, nC, status, nC_off, 0, 0, 0]."""
    on_boxes, off_boxes = _boxes(on["gt_boxes"], "online boxes"), _boxes(off["gt_boxes"], "offline boxes")
    on_cls, off_cls = _i64c(on["gt_classes"], "online classes"), _i64c(off["gt_classes"], "offline classes")
    on_s, off_s = _f32c(on["scores"], "online scores"), _f32c(off["scores"], "offline scores")
    on_p, off_p = _f32c(on["probs"], "online probs"), _f32c(off["probs"], "offline probs")
    nc, nd = on_boxes.shape[0], off_boxes.shape[0]
    k1 = int(on_p.shape[1])
    dev = on_boxes.device
    cap = nc * nd + nc + nd
    i32 = lambda n: torch.empty((max(n, 1),), dtype=torch.int32, device=dev)
    f32 = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
    i64 = lambda n: torch.empty((n,), dtype=torch.int64, device=dev)
    c_on, c_off = i32(nc + nd), i32(nc + nd)
    ws = _workspace(lib.coin_match_abc_workspace_bytes(nc, nd), dev)
    idx = {}
    for tag in tags:
        idx[tag] = {"a_on": i32(cap), "a_off": i32(cap), "a_box": f32(max(cap, 1), 4),
                    "counts": torch.zeros((8,), dtype=torch.int32, device=dev)}
        if tag == "RCNN":
            idx[tag].update({"b_on": i32(cap), "b_off": i32(cap), "b_box": f32(max(cap, 1), 4)})
    nd_ptr = _ptr(_count(nd_dev))
    if len(tags) == 2:
        r, p = idx["RCNN"], idx["RPN"]
        check(lib.coin_match_abc_both_dev(_ptr(on_boxes), _ptr(on_cls), _ptr(on_s), nc, _ptr(off_boxes), _ptr(off_cls),
                                          _ptr(off_s), nd, nd_ptr, float(iou_thr), float(weight_for_box_a), cap,
                                          _ptr(r["a_on"]), _ptr(r["a_off"]), _ptr(r["a_box"]), _ptr(r["b_on"]),
                                          _ptr(r["b_off"]), _ptr(r["b_box"]), _ptr(p["a_on"]), _ptr(p["a_off"]),
                                          _ptr(p["a_box"]), _ptr(c_on), _ptr(c_off), _ptr(r["counts"]), _ptr(p["counts"]),
                                          _ptr(ws), ws.numel(), _stream()))
    else:
        tag = tags[0]
        t = idx[tag]
        code = {"RCNN": _lib.TAG_RCNN, "RPN": _lib.TAG_RPN}[tag]
        check(lib.coin_match_abc_dev(_ptr(on_boxes), _ptr(on_cls), _ptr(on_s), nc, _ptr(off_boxes), _ptr(off_cls),
                                     _ptr(off_s), nd, nd_ptr, code, float(iou_thr), float(weight_for_box_a), cap,
                                     _ptr(t["a_on"]), _ptr(t["a_off"]), _ptr(t["a_box"]), _ptr(t.get("b_on")),
                                     _ptr(t.get("b_off")), _ptr(t.get("b_box")), _ptr(c_on), _ptr(c_off),
                                     _ptr(t["counts"]), _ptr(ws), ws.numel(), _stream()))

    def pseudo(n, boxes, split):
        d = {"gt_boxes": boxes}
        if split:
            d["gt_classes_offline"], d["gt_classes_online"] = i64(n), i64(n)
        else:
            d["gt_classes"] = i64(n)
        d["gt_scores_online"], d["gt_scores_offline"] = f32(n), f32(n)
        d["gt_probs_online"], d["gt_probs_offline"] = f32(n, k1), f32(n, k1)
        return d

    def dets_struct(boxes, cls, s, p):
        st = _lib.CoinDets()
        st.boxes, st.classes, st.scores, st.probs = boxes.data_ptr(), cls.data_ptr(), s.data_ptr(), p.data_ptr()
        return st

    def pseudo_struct(d, kind):
        st = _lib.CoinPseudo()
        st.boxes = d["gt_boxes"].data_ptr()
        if kind == "C":
            st.classes, st.scores_online, st.probs_online = (d["gt_classes"].data_ptr(), d["gt_scores"].data_ptr(),
                                                             d["gt_probs"].data_ptr())
            return st
        st.classes = (d["gt_classes_offline"] if kind == "B" else d["gt_classes"]).data_ptr()
        if kind == "B":
            st.classes_online = d["gt_classes_online"].data_ptr()
        st.scores_online, st.scores_offline = d["gt_scores_online"].data_ptr(), d["gt_scores_offline"].data_ptr()
        st.probs_online, st.probs_offline = d["gt_probs_online"].data_ptr(), d["gt_probs_offline"].data_ptr()
        return st

    on_st, off_st = dets_struct(on_boxes, on_cls, on_s, on_p), dets_struct(off_boxes, off_cls, off_s, off_p)
    out = {}
    for tag in tags:
        t = idx[tag]
        code = {"RCNN": _lib.TAG_RCNN, "RPN": _lib.TAG_RPN}[tag]
        a = pseudo(max(cap, 1), t["a_box"], False)
        b = pseudo(max(cap, 1), t["b_box"], True) if tag == "RCNN" else None
        c = {"gt_boxes": f32(max(nc + nd, 1), 4), "gt_classes": i64(max(nc + nd, 1)), "gt_scores": f32(max(nc + nd, 1)),
             "gt_probs": f32(max(nc + nd, 1), k1)}
        a_st, c_st = pseudo_struct(a, "A"), pseudo_struct(c, "C")
        b_st = pseudo_struct(b, "B") if b is not None else None
        check(lib.coin_abc_pack(ctypes.byref(on_st), nc, ctypes.byref(off_st), nd, nd_ptr, k1, code,
                                _ptr(t["a_on"]), _ptr(t["a_off"]), _ptr(t.get("b_on")), _ptr(t.get("b_off")), _ptr(c_on),
                                _ptr(c_off), _ptr(t["counts"]), ctypes.byref(a_st),
                                ctypes.byref(b_st) if b_st is not None else None, ctypes.byref(c_st), cap, _stream()))
        out[tag] = (a, b, c, t["counts"])
    return out


# ------------------------------------------------------------------------------------------------
# sampling and loss-side reductions
# ------------------------------------------------------------------------------------------------
def proposal_classes(matched_idxs: torch.Tensor, matched_labels: torch.Tensor, gt_classes: torch.Tensor, num_classes: int,
                     m_dev: Optional[torch.Tensor] = None, n_gt_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    """detectron2 ROIHeads._sample_proposals, first half: the class of every proposal (num_classes = background, -1 = ignore)."""
    idx = _i64c(matched_idxs, "matched_idxs")
    lab = _cuda(matched_labels, "matched_labels").to(torch.int8).contiguous()
    gt = _i64c(gt_classes, "gt_classes")
    out = torch.empty_like(idx)
    check(lib.coin_proposal_classes(_ptr(idx), _ptr(lab), _ptr(gt), gt.numel(), _ptr(None if n_gt_dev is None else _count(n_gt_dev)),
                                    idx.numel(), _ptr(None if m_dev is None else _count(m_dev)), int(num_classes), _ptr(out),
                                    _stream()))
    return out


def subsample_labels(labels: torch.Tensor, num_samples: int, positive_fraction: float, bg_label: int,
                     perms: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, seed: int = 0, offset: int = 0,
                     m_dev: Optional[torch.Tensor] = None, count_only: bool = False, sync: bool = True):
    """detectron2 subsample_labels. perms=(perm_pos, perm_neg): replay of the reference's two torch.randperm draws;
    None: the device generator (Philox4x32-10 keyed by seed / offset). Returns (pos_idx, neg_idx) - or, with sync=False,
    the capacity buffers and the device counts [num_pos, num_neg, P, N, status]; count_only returns (P, N)."""
    _cuda(labels, "labels")
    if labels.dtype not in (torch.int64, torch.int8):
        raise TypeError("coin_b200: labels must be int64 or int8")
    labels = labels.contiguous()
    m = labels.numel()
    dev = labels.device
    num_pos_target = int(num_samples * positive_fraction)
    pos = torch.empty((max(num_samples, 1),), dtype=torch.int64, device=dev)
    neg = torch.empty((max(num_samples, 1),), dtype=torch.int64, device=dev)
    counts = torch.zeros((8,), dtype=torch.int32, device=dev)
    ws = _workspace(lib.coin_subsample_labels_workspace_bytes(m), dev)
    pp = pn = None
    if perms is not None:
        # (an empty permutation has no storage: hand the library a valid pointer anyway, NULL means "device generator")
        pp, pn = (_i64c(p, "perm") if p.numel() else torch.zeros((1,), dtype=torch.int64, device=dev) for p in perms)
    check(lib.coin_subsample_labels(_ptr(labels), int(labels.dtype == torch.int8), m,
                                    _ptr(None if m_dev is None else _count(m_dev)), int(num_samples), num_pos_target,
                                    int(bg_label), _ptr(pp), _ptr(pn), int(seed) & (2 ** 64 - 1), int(offset) & (2 ** 64 - 1),
                                    int(bool(count_only)), _ptr(pos), _ptr(neg), _ptr(counts), _ptr(ws), ws.numel(), _stream()))
    if count_only:
        c = counts.tolist()
        return c[2], c[3]
    if not sync:
        return pos, neg, counts
    c = counts.tolist()
    if c[4]:
        raise ValueError("coin_b200: subsample_labels: a permutation entry is out of range")
    return pos[: c[0]], neg[: c[1]]


def rpn_teacher_probs(gt_probs: torch.Tensor, matched_idxs: torch.Tensor, nc_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    """rpn.py:95-98: gt_probs[:, :-1].sum(1)[matched_idxs]; zeros when there are no C boxes."""
    idx = _i64c(matched_idxs, "matched_idxs")
    out = torch.empty(idx.shape, dtype=torch.float32, device=idx.device)
    gp = _f32c(gt_probs, "gt_probs") if gt_probs is not None and gt_probs.numel() else None
    nc = 0 if gp is None else gp.shape[0]
    k1 = 1 if gp is None else int(gp.shape[1])
    check(lib.coin_rpn_teacher_probs(_ptr(gp), nc, _ptr(None if nc_dev is None else _count(nc_dev)), k1, _ptr(idx), idx.numel(),
                                     _ptr(out), _stream()))
    return out


def kl_distill_roi_fwd(scores: torch.Tensor, gt_probs: torch.Tensor, n_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    scores, gt_probs = _f32c(scores, "scores"), _f32c(gt_probs, "gt_probs")
    if scores.shape != gt_probs.shape or scores.dim() != 2:
        raise ValueError("coin_b200: scores and gt_probs must both be [n, K+1]")
    loss = torch.empty((), dtype=torch.float32, device=scores.device)
    ws = _workspace(lib.coin_kl_workspace_bytes(), scores.device)
    check(lib.coin_kl_distill_roi_fwd(_ptr(scores), _ptr(gt_probs), scores.shape[0],
                                      _ptr(None if n_dev is None else _count(n_dev)), int(scores.shape[1]), _ptr(loss), _ptr(ws),
                                      _stream()))
    return loss


def kl_distill_roi_bwd(scores, gt_probs, grad_loss, n_dev=None) -> torch.Tensor:
    scores, gt_probs = _f32c(scores, "scores"), _f32c(gt_probs, "gt_probs")
    g = torch.empty_like(scores)
    go = _f32c(grad_loss, "grad_loss").reshape(1)
    check(lib.coin_kl_distill_roi_bwd(_ptr(scores), _ptr(gt_probs), scores.shape[0],
                                      _ptr(None if n_dev is None else _count(n_dev)), int(scores.shape[1]), _ptr(go), _ptr(g),
                                      _stream()))
    return g


def kl_distill_rpn_fwd(logits: torch.Tensor, distillation_labels: torch.Tensor, teacher_probs: torch.Tensor):
    """Returns (loss, n_valid device int32)."""
    logits = _f32c(logits, "logits").reshape(-1)
    labels = _cuda(distillation_labels, "distillation_labels").to(torch.int8).contiguous().reshape(-1)
    teacher = _f32c(teacher_probs, "teacher_probs").reshape(-1)
    if not (logits.numel() == labels.numel() == teacher.numel()):
        raise ValueError("coin_b200: logits, labels and teacher_probs disagree in length")
    loss = torch.empty((), dtype=torch.float32, device=logits.device)
    n_valid = torch.zeros((1,), dtype=torch.int32, device=logits.device)
    ws = _workspace(lib.coin_kl_workspace_bytes(), logits.device)
    check(lib.coin_kl_distill_rpn_fwd(_ptr(logits), _ptr(labels), _ptr(teacher), logits.numel(), _ptr(loss), _ptr(n_valid),
                                      _ptr(ws), _stream()))
    return loss, n_valid


def kl_distill_rpn_bwd(logits, distillation_labels, teacher_probs, n_valid, grad_loss) -> torch.Tensor:
    shape = logits.shape
    logits = _f32c(logits, "logits").reshape(-1)
    labels = _cuda(distillation_labels, "distillation_labels").to(torch.int8).contiguous().reshape(-1)
    teacher = _f32c(teacher_probs, "teacher_probs").reshape(-1)
    g = torch.empty_like(logits)
    go = _f32c(grad_loss, "grad_loss").reshape(1)
    check(lib.coin_kl_distill_rpn_bwd(_ptr(logits), _ptr(labels), _ptr(teacher), logits.numel(), _ptr(_count(n_valid)), _ptr(go),
                                      _ptr(g), _stream()))
    return g.reshape(shape)

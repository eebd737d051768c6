"""ctypes loader for oracle/_build/liboracle.so (TEST INFRASTRUCTURE ONLY)."""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "scalar_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _f32(t):
    assert t.device.type == "cpu"
    return t.detach().to(torch.float32).contiguous()


def roi_align_fwd(x, rois, spatial_scale, ph, pw, sampling_ratio, aligned):
    x, rois = _f32(x), _f32(rois)
    n, c, h, w = x.shape
    k = rois.shape[0]
    out = torch.empty((k, c, ph, pw), dtype=torch.float32)
    lib().orc_roi_align_fwd(_p(x), _p(rois), _p(out), n, c, h, w, k, ph, pw,
                            ctypes.c_float(spatial_scale), int(sampling_ratio), int(bool(aligned)))
    return out


def roi_align_bwd(grad, rois, spatial_scale, ph, pw, n, c, h, w, sampling_ratio, aligned):
    grad, rois = _f32(grad), _f32(rois)
    k = rois.shape[0]
    gin = torch.empty((n, c, h, w), dtype=torch.float32)
    lib().orc_roi_align_bwd(_p(grad), _p(rois), _p(gin), n, c, h, w, k, ph, pw,
                            ctypes.c_float(spatial_scale), int(sampling_ratio), int(bool(aligned)))
    return gin


def pairwise_iou(b1, b2):
    b1, b2 = _f32(b1), _f32(b2)
    out = torch.empty((b1.shape[0], b2.shape[0]), dtype=torch.float32)
    lib().orc_pairwise_iou(_p(b1), b1.shape[0], _p(b2), b2.shape[0], _p(out))
    return out


def nms(boxes, scores, thr):
    boxes, scores = _f32(boxes), _f32(scores)
    n = boxes.shape[0]
    keep = torch.empty((max(n, 1),), dtype=torch.int64)
    nk = ctypes.c_int64(0)
    lib().orc_nms(_p(boxes), _p(scores), ctypes.c_int64(n), ctypes.c_double(thr), _p(keep),
                  ctypes.byref(nk))
    return keep[: nk.value].clone()

"""ctypes binding of libcoinops.so (the C ABI declared in include/coinops.h).

There is no fallback: if the shared library is missing or a symbol cannot be resolved, importing
this module raises. Every call checks the return code and raises with ``coin_last_error()``.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libcoinops.so")

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_CAPACITY = 0, 1, 2, 3, 4
F32, F16 = 0, 1
NMS_PLAIN, NMS_TRICK, NMS_VANILLA, NMS_AUTO = 0, 1, 2, 3
SCORE_PROBEN, SCORE_AVG, SCORE_MAX = 0, 1, 2
BOX_SAVG, BOX_AVG, BOX_MAX = 0, 1, 2
TAG_RCNN, TAG_RPN = 0, 1
MAX_LEVELS = 8
ABC_MAX = 1024


class CoinLevel(Structure):
    _fields_ = [("feat_nhwc", c_void_p), ("H", c_int), ("W", c_int), ("spatial_scale", c_float)]


class CoinSeg(Structure):
    _fields_ = [("ptr", c_void_p), ("count_dev", c_void_p), ("count", c_int64), ("prefix", c_float)]


class CoinDets(Structure):
    _fields_ = [("boxes", c_void_p), ("classes", c_void_p), ("scores", c_void_p), ("probs", c_void_p)]


class CoinPseudo(Structure):
    _fields_ = [("boxes", c_void_p), ("classes", c_void_p), ("classes_online", c_void_p),
                ("scores_online", c_void_p), ("scores_offline", c_void_p), ("probs_online", c_void_p),
                ("probs_offline", c_void_p)]


P = c_void_p
_SIGNATURES = {
    # name: (restype, [argtypes])
    "coin_last_error": (c_char_p, []),
    "coin_version": (c_int, []),
    "coin_launch_count": (ctypes.c_longlong, []),
    "coin_set_option": (c_int, [c_char_p, c_int]),
    "coin_unset_option": (c_int, [c_char_p]),
    "coin_get_option": (c_int, [c_char_p, c_int]),
    "coin_nchw_to_nhwc_f32": (c_int, [P, c_int, P, c_int, c_int, c_int, c_int, P]),
    "coin_nhwc_f32_to_nchw": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "coin_roi_align_fwd": (c_int, [POINTER(CoinLevel), c_int, P, P, P, c_int, c_int, c_int, c_int, c_int,
                                   c_int, c_int, P]),
    "coin_roi_align_bwd": (c_int, [POINTER(CoinLevel), c_int, P, P, P, c_int, c_int, c_int, c_int, c_int,
                                   c_int, c_int, P]),
    "coin_roi_launch_order": (c_int, [P, c_int, P, c_int, c_int, P, P]),
    "coin_roi_launch_plan": (c_int, [P, c_int, P, c_int, c_int, c_float, c_float, c_int, P, P, P, P]),
    "coin_roi_split_by_area": (c_int, [P, c_int, P, c_float, c_float, c_int, P, P, P, P, P]),
    "coin_roi_align_fwd_ord": (c_int, [POINTER(CoinLevel), c_int, P, P, P, c_int, c_int, c_int, c_int, c_int,
                                       c_int, c_int, P, P, P]),
    "coin_roi_align_bwd_ord": (c_int, [POINTER(CoinLevel), c_int, P, P, P, c_int, c_int, c_int, c_int, c_int,
                                       c_int, c_int, P, P, P]),
    "coin_roi_pooler_levels": (c_int, [P, c_int64, c_int, c_int, c_int, c_int, P, P]),
    "coin_apply_deltas": (c_int, [P, P, P, c_int64, c_int, c_float, c_float, c_float, c_float, c_float,
                                  c_int, c_float, c_float, P]),
    "coin_get_deltas": (c_int, [P, P, P, c_int64, c_float, c_float, c_float, c_float, P, P]),
    "coin_boxes_clip": (c_int, [P, c_int64, c_float, c_float, P]),
    "coin_boxes_scale_flip": (c_int, [P, P, c_int64, c_float, c_float, c_int, c_float, c_float, P]),
    "coin_boxes_cxcywh_to_xyxy": (c_int, [P, P, c_int64, c_float, c_float, c_int, P]),
    "coin_pairwise_iou": (c_int, [P, c_int64, P, c_int64, P, P]),
    "coin_matcher": (c_int, [P, c_int64, c_int64, POINTER(c_float), c_int, POINTER(ctypes.c_int8), c_int,
                             P, P, P, P, P]),
    "coin_iou_match_workspace_floats": (c_size_t, [c_int64, c_int64]),
    "coin_iou_match": (c_int, [P, c_int64, P, c_int64, POINTER(c_float), c_int, POINTER(ctypes.c_int8),
                               c_int, P, P, P, P, P]),
    "coin_relabel_roi": (c_int, [P, P, c_int64, c_int64, c_int64, P]),
    "coin_relabel_rpn": (c_int, [P, P, c_int64, c_int64, c_int64, P, P, P]),
    "coin_iou_pairs_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "coin_iou_pairs_ge": (c_int, [P, c_int64, P, c_int64, c_float, P, P, c_int64, P, c_size_t, P]),
    "coin_nms_workspace_bytes": (c_size_t, [c_int64]),
    "coin_batched_nms": (c_int, [P, P, P, c_int64, c_double, c_int, c_int64, P, P, P, c_size_t, P]),
    "coin_fusion_nms_workspace_bytes": (c_size_t, [c_int64, c_int]),
    "coin_fusion_nms": (c_int, [P, P, P, c_int64, c_int, c_float, c_int, c_int, c_int, P, P, P, P, P, P, P,
                                P, c_size_t, P]),
    "coin_det_postprocess_workspace_bytes": (c_size_t, [c_int64, c_int]),
    "coin_det_postprocess": (c_int, [P, P, c_int64, c_int, c_int, c_float, c_float, c_float, c_double,
                                     c_int64, c_int64, P, P, P, P, P, P, P, c_size_t, P]),
    "coin_match_abc_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "coin_match_abc": (c_int, [P, P, P, c_int64, P, P, P, c_int64, c_int, c_float, c_float, c_int64,
                               P, P, P, P, P, P, P, P, P, P, c_size_t, P]),
    # sync-free (device-count) variants
    "coin_roi_align_fwd_dev": (c_int, [POINTER(CoinLevel), c_int, P, P, P, c_int, c_int, c_int, c_int, c_int,
                                       c_int, c_int, P, P]),
    "coin_iou_match_dev": (c_int, [P, c_int64, P, P, c_int64, P, POINTER(c_float), c_int, POINTER(ctypes.c_int8),
                                   c_int, P, P, P, P, P]),
    "coin_relabel_roi_dev": (c_int, [P, P, c_int64, P, P, P, P, P]),
    "coin_relabel_rpn_dev": (c_int, [P, P, c_int64, P, P, P, P, P]),
    "coin_match_abc_dev": (c_int, [P, P, P, c_int64, P, P, P, c_int64, P, c_int, c_float, c_float, c_int64,
                                   P, P, P, P, P, P, P, P, P, P, c_size_t, P]),
    "coin_match_abc_both_dev": (c_int, [P, P, P, c_int64, P, P, P, c_int64, P, c_float, c_float, c_int64,
                                        P, P, P, P, P, P, P, P, P, P, P, P, P, P, c_size_t, P]),
    "coin_rpn_proposals_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "coin_rpn_proposals": (c_int, [P, P, P, c_int64, c_int64, c_int64, c_double, c_float, c_float, c_float, c_float,
                                   c_float, c_float, c_float, c_float, P, P, P, P, P, c_size_t, P]),
    "coin_rpn_proposals_grid": (c_int, [POINTER(c_float), c_int, c_int, c_int, c_float, c_float, P, P, c_int64, c_int64, c_double,
                                        c_float, c_float, c_float, c_float, c_float, c_float, c_float, c_float, P, P, P, P, P,
                                        c_size_t, P]),
    "coin_concat_rows": (c_int, [POINTER(CoinSeg), c_int, c_int, c_int, P, c_int64, P, P]),
    "coin_pack_rows": (c_int, [P, c_int, P, P, c_int64, P, P]),
    "coin_abc_pack": (c_int, [POINTER(CoinDets), c_int64, POINTER(CoinDets), c_int64, P, c_int, c_int,
                              P, P, P, P, P, P, P, POINTER(CoinPseudo), POINTER(CoinPseudo), POINTER(CoinPseudo),
                              c_int64, P]),
    # sampling and loss-side reductions
    "coin_proposal_classes": (c_int, [P, P, P, c_int64, P, c_int64, P, c_int64, P, P]),
    "coin_subsample_labels_workspace_bytes": (c_size_t, [c_int64]),
    "coin_subsample_labels": (c_int, [P, c_int, c_int64, P, c_int, c_int, c_int64, P, P, ctypes.c_uint64, ctypes.c_uint64,
                                      c_int, P, P, P, P, c_size_t, P]),
    "coin_rpn_teacher_probs": (c_int, [P, c_int64, P, c_int, P, c_int64, P, P]),
    "coin_kl_workspace_bytes": (c_size_t, []),
    "coin_kl_distill_roi_fwd": (c_int, [P, P, c_int64, P, c_int, P, P, P]),
    "coin_kl_distill_roi_bwd": (c_int, [P, P, c_int64, P, c_int, P, P, P]),
    "coin_kl_distill_rpn_fwd": (c_int, [P, P, P, c_int64, P, P, P, P]),
    "coin_kl_distill_rpn_bwd": (c_int, [P, P, P, c_int64, P, P, P, P]),
    # evaluation
    "coin_argsort_desc_workspace_bytes": (c_size_t, [c_int64]),
    "coin_argsort_desc": (c_int, [P, c_int64, P, P, c_size_t, P]),
    "coin_voc_match_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "coin_voc_match": (c_int, [P, P, P, c_int64, P, P, P, c_int64, c_double, P, P, P, c_size_t, P]),
    # peer-memory all-reduce
    "coin_p2p_all_reduce": (c_int, [POINTER(c_void_p), POINTER(c_void_p), c_int, c_int, c_int64, c_int64, ctypes.c_uint32, P, c_int, P]),
}

EXPORTS = tuple(_SIGNATURES.keys())

if not os.path.exists(SO_PATH):
    raise ImportError(
        f"{SO_PATH} is missing: build it with `python coin_b200/build.py` (nvcc, sm_100a). "
        "coin_b200 has no CPU or PyTorch fallback.")

lib = ctypes.CDLL(SO_PATH)
for _name, (_res, _args) in _SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here = the library does not export the declared ABI
    _fn.restype = _res
    _fn.argtypes = _args


def set_option(name: str, value) -> None:
    """Process-wide mode / tuning switch of the library (include/coinops.h: coin_set_option); None unsets."""
    if value is None:
        check(lib.coin_unset_option(name.encode()))
    else:
        check(lib.coin_set_option(name.encode(), int(value)))


def get_option(name: str, default: int = 0) -> int:
    return int(lib.coin_get_option(name.encode(), int(default)))


class options:
    """Context manager: ``with _lib.options(COIN_ROI_EXACT=1): ...`` (restores the previous state)."""

    def __init__(self, **kw):
        self.kw, self.old = kw, {}

    def __enter__(self):
        for k, v in self.kw.items():
            self.old[k] = get_option(k, -(2 ** 31))
            set_option(k, v)
        return self

    def __exit__(self, *exc):
        for k, v in self.old.items():
            set_option(k, None if v == -(2 ** 31) else v)


class CoinError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libcoinops error {code}: {message}")
        self.code = code


def check(rc: int) -> None:
    if rc != OK:
        msg = lib.coin_last_error().decode("utf-8", "replace")
        if rc == ERR_INVALID:
            raise ValueError(f"libcoinops: {msg}")
        raise CoinError(rc, msg)

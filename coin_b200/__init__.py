"""coin_b200 -- B200-native (sm_100a) RoI / box-decode / IoU-match / NMS path of Flashkong/COIN.

Importing the package loads ``libcoinops.so`` (hand-written CUDA kernels behind a C ABI, see
include/coinops.h). There is no CPU, PyTorch-eager or Triton fallback: if the library is missing the
import fails, and CPU tensors are rejected by every operator.
"""
from . import _lib  # noqa: F401  (raises ImportError when the extension is not built)
from .layers import (Box2BoxTransform, Matcher, MyNMS, ROIAlign, ROIPooler, batched_nms, mynms, nms,
                     pairwise_iou)
from .structures import Boxes, Instances

__all__ = ["Box2BoxTransform", "Matcher", "MyNMS", "ROIAlign", "ROIPooler", "batched_nms", "mynms", "nms",
           "pairwise_iou", "Boxes", "Instances"]
__version__ = "0.1.0"

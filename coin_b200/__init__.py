"""coin_b200 -- B200-native (sm_100a) RoI / box-decode / IoU-match / NMS path of Flashkong/COIN.

Importing the package loads ``libcoinops.so`` (hand-written CUDA kernels behind a C ABI, see
include/coinops.h). There is no CPU, PyTorch-eager or Triton fallback: if the library is missing the
import fails, and CPU tensors are rejected by every operator.
"""
import os as _os

# The step runs ~10 concurrent streams (coin_b200/pipeline.py); the default of 8 hardware work queues
# would serialise some of them behind each other. Must be set before the CUDA context is created.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from . import _lib  # noqa: F401,E402  (raises ImportError when the extension is not built)
from .layers import (Box2BoxTransform, Matcher, MyNMS, ROIAlign, ROIPooler, batched_nms, mynms, nms,
                     pairwise_iou)
from .structures import Boxes, Instances
from .dispatch import is_patched, patch, unpatch

__all__ = ["Box2BoxTransform", "Matcher", "MyNMS", "ROIAlign", "ROIPooler", "batched_nms", "mynms", "nms",
           "pairwise_iou", "Boxes", "Instances", "patch", "unpatch", "is_patched"]
__version__ = "0.1.0"

"""detectron2.layers.batched_nms stand-in (restated from detectron2 0.5 layers/nms.py)."""
from oracle.d2_ref import batched_nms  # noqa: F401

"""detectron2.layers: batched_nms, nonzero_tuple, cat (restated from detectron2 0.5 layers/nms.py,
layers/wrappers.py)."""
from typing import List

import torch

from oracle.d2_ref import batched_nms  # noqa: F401  (the d2 wrapper over torchvision's CPU batched_nms)


def cat(tensors: List[torch.Tensor], dim: int = 0):
    assert isinstance(tensors, (list, tuple))
    if len(tensors) == 1:
        return tensors[0]
    return torch.cat(tensors, dim)


def nonzero_tuple(x):
    if x.dim() == 0:
        return x.unsqueeze(0).nonzero().unbind(1)
    return x.nonzero().unbind(1)

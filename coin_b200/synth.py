"""Seeded synthetic inputs of the shapes SURVEY.md section 8(d) prescribes.

Everything is generated on the CPU (``torch.Generator`` seeded with the reference's
``SEED: 2024``, configs/coin/GDINO/foggy.yaml:44) and copied to the device afterwards, so that the
CPU oracle and the CUDA path see identical bits. No dataset or checkpoint is read.
"""
import math
from dataclasses import dataclass
from typing import Dict, List, Tuple

import torch

SEED = 2024


@dataclass
class Shape:
    """One benchmark / parity configuration (names follow BASELINE.json ``configs``)."""
    name: str
    images: int
    height: int
    width: int
    classes: int
    rois: int            # RoIs per image fed to ROIAlign
    pooled: int          # ROIAlign output size
    proposals: int       # proposals per image labelled by IoU + Matcher
    teacher_rois: int    # RoIs per image decoded / NMS-ed by the teacher branch
    cloud: int = 100
    clip: int = 100
    channels: int = 1024
    stride: int = 16
    rpn_pre_nms: int = 6000
    rpn_post_nms: int = 1000

    @property
    def feat_hw(self) -> Tuple[int, int]:
        h, w = self.height, self.width
        for _ in range(4):  # CLIP ModifiedResNet: four floor-halvings to stride 16 (utils.py:219-238)
            h, w = h // 2, w // 2
        return h, w


SHAPES: Dict[str, Shape] = {
    # configs[0]: the reference's own CPU-runnable case
    "foggy_cpu": Shape("foggy_cpu", 2, 600, 1200, 8, 512, 7, 2000, 1000),
    # configs[1]: RoI-head forward/backward, Foggy-Cityscapes shape (the bench workload at N=1)
    "foggy_roi_head": Shape("foggy_roi_head", 3, 600, 1200, 8, 512, 14, 2000, 1000,
                            rpn_pre_nms=12000, rpn_post_nms=2000),
    # configs[3]: BDD100K shape, 7 classes, 2000 RoIs / image
    "bdd_2000": Shape("bdd_2000", 3, 600, 1067, 7, 2000, 14, 2000, 1000,
                      rpn_pre_nms=12000, rpn_post_nms=2000),
    # tiny shape for smoke tests
    "tiny": Shape("tiny", 2, 160, 320, 8, 48, 7, 200, 120, cloud=30, clip=30, channels=64,
                  rpn_pre_nms=600, rpn_post_nms=100),
}


def gen(seed: int = SEED) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def _log_uniform(g, n, lo, hi):
    return torch.exp(torch.rand(n, generator=g) * (math.log(hi) - math.log(lo)) + math.log(lo))


def random_boxes(g, n: int, height: int, width: int, lo: float = 16.0, hi: float = 500.0,
                 min_side: float = 4.0) -> torch.Tensor:
    """Boxes with log-uniform sides, centres uniform in the image, clipped, x2>x1 and y2>y1."""
    w = _log_uniform(g, n, lo, hi).clamp(max=float(width))
    h = _log_uniform(g, n, lo, hi).clamp(max=float(height))
    cx = torch.rand(n, generator=g) * width
    cy = torch.rand(n, generator=g) * height
    x1 = (cx - w / 2).clamp(0, width - min_side)
    y1 = (cy - h / 2).clamp(0, height - min_side)
    x2 = torch.maximum((cx + w / 2).clamp(0, width), x1 + min_side)
    y2 = torch.maximum((cy + h / 2).clamp(0, height), y1 + min_side)
    return torch.stack((x1, y1, x2, y2), dim=1).float()


def jitter(g, boxes: torch.Tensor, rel: float, height: int, width: int, abs_px: float = 0.0):
    wh = torch.stack((boxes[:, 2] - boxes[:, 0], boxes[:, 3] - boxes[:, 1]), dim=1).repeat(1, 2)
    out = boxes + torch.randn(boxes.shape, generator=g) * (rel * wh + abs_px)
    x1 = out[:, 0].clamp(0, width - 2.0)
    y1 = out[:, 1].clamp(0, height - 2.0)
    x2 = torch.maximum(out[:, 2].clamp(0, width), x1 + 2.0)
    y2 = torch.maximum(out[:, 3].clamp(0, height), y1 + 2.0)
    return torch.stack((x1, y1, x2, y2), dim=1).float()


def objects(g, shape: Shape, n: int = 40) -> torch.Tensor:
    return random_boxes(g, n, shape.height, shape.width)


def rois_for(g, shape: Shape, objs: torch.Tensor, n: int) -> torch.Tensor:
    """25 % jittered copies of objects (sigma = 10 % of the side) + 75 % random boxes."""
    n_obj = n // 4
    pick = torch.randint(0, objs.shape[0], (n_obj,), generator=g)
    a = jitter(g, objs[pick], 0.10, shape.height, shape.width)
    b = random_boxes(g, n - n_obj, shape.height, shape.width)
    out = torch.cat((a, b), dim=0)
    return out[torch.randperm(n, generator=g)]


def _probs(g, n, k1, bg_zero: bool):
    logits = 3.0 * torch.randn(n, k1, generator=g)
    if bg_zero:
        logits[:, -1] = -float("inf")
    return torch.softmax(logits, dim=1)


def cloud_dets(g, shape: Shape, objs: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Cloud (GDINO-like) detections: objects + N(0, 2 px) jitter; 20 % of them get a near-copy
    (IoU >= 0.95) with a different class (exercises online_boxes_merging); remainder random.
    probs have a zero background column (gdino.py:187-188); score = max prob, class = argmax."""
    k1 = shape.classes + 1
    n = shape.cloud
    n_obj = min(objs.shape[0], int(n * 0.6))
    base = jitter(g, objs[:n_obj], 0.0, shape.height, shape.width, abs_px=2.0)
    n_twin = min(int(n_obj * 0.2), n - n_obj)
    twin_src = torch.randperm(n_obj, generator=g)[:n_twin]
    twins = jitter(g, base[twin_src], 0.0, shape.height, shape.width, abs_px=0.25)
    rest = random_boxes(g, n - n_obj - n_twin, shape.height, shape.width)
    boxes = torch.cat((base, twins, rest), dim=0)
    probs = _probs(g, n, k1, bg_zero=True)
    cls = probs.argmax(1)
    # force the twins onto another class than their source
    for t, s in enumerate(twin_src.tolist()):
        j = n_obj + t
        if cls[j] == cls[s]:
            alt = int((cls[s] + 1) % shape.classes)
            row = probs[j].clone()
            row[alt], row[cls[j]] = probs[j, cls[j]], probs[j, alt]
            probs[j] = row
    cls = probs.argmax(1)
    return {"gt_boxes": boxes, "gt_classes": cls, "scores": probs.max(1)[0], "probs": probs}


def clip_dets(g, shape: Shape, objs: torch.Tensor, cloud: Dict[str, torch.Tensor]):
    """CLIP-detector detections: jittered objects carrying the cloud class with 20 % flips, 15 %
    exact duplicates with another class (exercises delete_duplicate_boxes), random extras;
    pairwise-distinct scores in (0.05, 1)."""
    k1 = shape.classes + 1
    n = shape.clip
    n_obj = min(objs.shape[0], int(n * 0.55))
    base = jitter(g, objs[:n_obj], 0.03, shape.height, shape.width)
    n_dup = min(int(n * 0.15), n_obj)
    dup_src = torch.randperm(n_obj, generator=g)[:n_dup]
    dups = base[dup_src].clone()
    rest = random_boxes(g, n - n_obj - n_dup, shape.height, shape.width)
    boxes = torch.cat((base, dups, rest), dim=0)
    cls = torch.randint(0, shape.classes, (n,), generator=g)
    cls[:n_obj] = cloud["gt_classes"][:n_obj]
    flip = torch.rand(n_obj, generator=g) < 0.2
    cls[:n_obj][flip] = (cls[:n_obj][flip] + 1 + torch.randint(0, shape.classes - 1, (int(flip.sum()),),
                                                            generator=g)) % shape.classes
    cls[n_obj:n_obj + n_dup] = (cls[dup_src] + 1 + torch.arange(n_dup) % (shape.classes - 1)) % shape.classes
    scores = 0.05 + 0.9 * torch.rand(n, generator=g) + torch.arange(n, dtype=torch.float32) * 2.0 ** -20
    probs = torch.full((n, k1), 0.0)
    probs[torch.arange(n), cls] = scores
    spread = (1.0 - scores) / (k1 - 1)
    probs = probs + spread[:, None] * (torch.arange(k1)[None, :] != cls[:, None])
    perm = torch.randperm(n, generator=g)
    return {"gt_boxes": boxes[perm], "gt_classes": cls[perm], "scores": scores[perm], "probs": probs[perm]}


def deltas_scores(g, r: int, k1: int):
    """Class-agnostic regression deltas (a few rows at +-10 to hit scale_clamp) and class logits."""
    d = 0.1 * torch.randn(r, 4, generator=g)
    if r >= 8:
        rows = torch.randint(0, r, (max(r // 128, 2),), generator=g)
        d[rows] = torch.where(torch.rand(rows.numel(), 4, generator=g) < 0.5, 10.0, -10.0)
    logits = 2.0 * torch.randn(r, k1, generator=g)
    return d, logits


def features(g, shape: Shape, dtype=torch.float32) -> torch.Tensor:
    h, w = shape.feat_hw
    return torch.randn(shape.images, shape.channels, h, w, generator=g).to(dtype)


def image_batch(shape: Shape, seed: int = SEED) -> Dict[str, object]:
    """All per-image inputs of one RoI-path step for ``shape`` (lists are per image)."""
    g = gen(seed)
    k1 = shape.classes + 1
    out: Dict[str, object] = {"shape": shape, "features": features(g, shape)}
    per_image: List[Dict[str, torch.Tensor]] = []
    for _ in range(shape.images):
        objs = objects(g, shape)
        cloud = cloud_dets(g, shape, objs)
        clip = clip_dets(g, shape, objs, cloud)
        teacher_rois = rois_for(g, shape, objs, shape.teacher_rois)
        t_deltas, t_logits = deltas_scores(g, shape.teacher_rois, k1)
        proposals = rois_for(g, shape, objs, shape.proposals)
        rpn_boxes = random_boxes(g, shape.rpn_pre_nms, shape.height, shape.width)
        rpn_scores = torch.randn(shape.rpn_pre_nms, generator=g)
        per_image.append({
            "objects": objs, "cloud": cloud, "clip": clip,
            "teacher_rois": teacher_rois, "teacher_deltas": t_deltas,
            "teacher_probs": torch.softmax(t_logits, dim=1),
            "proposals": proposals, "rois": rois_for(g, shape, objs, shape.rois),
            "rpn_boxes": rpn_boxes, "rpn_scores": rpn_scores,
        })
    out["images"] = per_image
    return out

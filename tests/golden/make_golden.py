"""Generates the golden fixtures of tests/golden/ (run in the BUILD container, where /root/reference
and torchvision's CPU operators exist):

    python tests/golden/make_golden.py

* fusion_nms_<method>.pt : outputs of the reference's OWN coin/layers/nms.py (loaded unmodified from
  /root/reference through oracle/ref_loader.py and the detectron2 stand-in in oracle/d2_shim) on seeded cloud-like detections.
* tv_ops.pt              : torchvision 0.26 CPU roi_align (fwd + bwd) / nms / batched_nms / box_iou
  on small seeded inputs, including adversarial RoIs (outside the map, sub-bin, whole map, inverted).
The GPU box has no /root/reference; tests only read the .pt files.
"""
import os
import sys

import torch
import torchvision

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)

from coin_b200 import synth  # noqa: E402


def load_reference_nms():
    """The reference's coin/layers/nms.py, loaded unmodified by path behind the detectron2 stand-in."""
    from oracle import ref_loader
    return ref_loader.load("coin.layers.nms")


def cloud_like(seed, n, k, height=600, width=1200):
    """Clustered detections with argmax(prob) == label and pairwise-distinct scores."""
    g = synth.gen(seed)
    n_obj = max(n // 4, 1)
    objs = synth.random_boxes(g, n_obj, height, width)
    pick = torch.randint(0, n_obj, (n,), generator=g)
    boxes = synth.jitter(g, objs[pick], 0.06, height, width)
    obj_cls = torch.randint(0, k, (n_obj,), generator=g)
    labels = obj_cls[pick].clone()
    flip = torch.rand(n, generator=g) < 0.15
    labels[flip] = torch.randint(0, k, (int(flip.sum()),), generator=g)
    logits = torch.randn(n, k + 1, generator=g)
    logits[:, -1] = -float("inf")
    logits[torch.arange(n), labels] = logits.max(1)[0] + 0.5 + 2.0 * torch.rand(n, generator=g)
    probs = torch.softmax(logits, dim=1)
    assert bool((probs.argmax(1) == labels).all())
    scores = probs.max(1)[0]
    assert scores.unique().numel() == n
    return boxes, scores, probs, labels


def main():
    ref = load_reference_nms()
    for method in ("ms", "ma", "ps", "pa", "pm", "as", "aa", "am", "nms", "mm"):
        cases = []
        for seed, n, k in ((1, 100, 8), (2, 37, 8), (3, 250, 20), (4, 1, 8), (5, 2, 3)):
            boxes, scores, probs, labels = cloud_like(seed, n, k)
            mynms = ref.MyNMS(method)
            keep, ob, os_, op, ol = mynms.nms(boxes.clone(), scores.clone(), probs.clone(), labels.clone(), 0.6)
            cases.append({"boxes": boxes, "scores": scores, "probs": probs, "labels": labels, "thr": 0.6,
                          "keep": keep, "out_boxes": ob, "out_scores": os_, "out_probs": op, "out_classes": ol})
        torch.save({"method": method, "cases": cases, "source": "/root/reference/coin/layers/nms.py"},
                   os.path.join(HERE, f"fusion_nms_{method}.pt"))

    # torchvision CPU operators
    g = synth.gen(7)
    n, c, h, w = 2, 8, 13, 19
    x = torch.randn(n, c, h, w, generator=g)
    rois = torch.tensor([
        [0, 10.0, 12.0, 150.0, 100.0], [1, -40.0, -30.0, 60.0, 50.0], [0, 0.0, 0.0, 304.0, 208.0],
        [1, 100.0, 100.0, 101.0, 101.5], [0, 280.0, 190.0, 400.0, 300.0], [1, 150.0, 80.0, 120.0, 60.0],
        [0, 500.0, 500.0, 600.0, 600.0], [1, 33.3, 47.7, 211.9, 160.1], [0, 16.0, 16.0, 32.0, 32.0]])
    tv = {"x": x, "rois": rois, "roi_align": []}
    for (ph, pw) in ((7, 7), (14, 14), (3, 5)):
        for sr in (0, 2):
            for aligned in (True, False):
                xx = x.clone().requires_grad_(True)
                out = torchvision.ops.roi_align(xx, rois, (ph, pw), 1.0 / 16, sr, aligned)
                go = torch.randn(out.shape, generator=g)
                out.backward(go)
                tv["roi_align"].append({"ph": ph, "pw": pw, "sr": sr, "aligned": aligned, "out": out.detach(),
                                        "grad_out": go, "grad_in": xx.grad.clone()})
    boxes = synth.jitter(g, synth.random_boxes(g, 12, 300, 400).repeat(25, 1), 0.08, 300, 400)
    scores = torch.rand(300, generator=g)
    scores[17] = scores[3]; scores[44] = scores[3]          # ties
    boxes[17] = boxes[3]                                     # identical box, identical score
    idxs = torch.randint(0, 5, (300,), generator=g)
    tv["nms"] = {"boxes": boxes, "scores": scores, "idxs": idxs,
                 "keep_0.5": torchvision.ops.nms(boxes, scores, 0.5),
                 "keep_0.7": torchvision.ops.nms(boxes, scores, 0.7),
                 "batched_keep_0.5": torchvision.ops.batched_nms(boxes, scores, idxs, 0.5),
                 "iou": torchvision.ops.box_iou(boxes[:40], boxes[40:110])}
    big = synth.jitter(g, synth.random_boxes(g, 60, 600, 1200).repeat(25, 1), 0.1, 600, 1200)
    bscores = torch.rand(1500, generator=g)
    bidx = torch.randint(0, 8, (1500,), generator=g)
    tv["nms_big"] = {"boxes": big, "scores": bscores, "idxs": bidx,
                     "batched_keep_0.5": torchvision.ops.batched_nms(big, bscores, bidx, 0.5)}  # numel>4000: vanilla
    tv["versions"] = {"torch": torch.__version__, "torchvision": torchvision.__version__}
    torch.save(tv, os.path.join(HERE, "tv_ops.pt"))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()

"""Freezes outputs of the reference's UNMODIFIED VOC evaluator (coin/evaluation/cloud_pascal_voc_evaluation.py:205-319,
voc_eval + voc_ap + parse_rec) executed here through oracle/ref_loader.py on synthetic annotation / detection FILES laid
out as the evaluator expects them (VOC XML per image, image-set list, one detection text file per class).

    python tests/golden/make_golden_voc.py        ->  tests/golden/voc_eval_ref.pt

The fixture stores the arrays behind the files (so the tests never touch /root/reference) and rec / prec / ap per class."""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

CLASSES = ("person", "car", "bus")


def write_xml(path, objs):
    rows = "".join(f"<object><name>{n}</name><difficult>{int(d)}</difficult><bndbox><xmin>{b[0]}</xmin><ymin>{b[1]}</ymin>"
                   f"<xmax>{b[2]}</xmax><ymax>{b[3]}</ymax></bndbox></object>" for n, d, b in objs)
    with open(path, "w") as f:
        f.write(f"<annotation>{rows}</annotation>")


def main():
    mod = ref_loader.load("coin.evaluation.cloud_pascal_voc_evaluation")
    rng = np.random.RandomState(2024)
    cases = []
    for case_id, (n_img, n_det, empty_class) in enumerate(((12, 400, None), (3, 40, "bus"), (30, 900, None))):
        with tempfile.TemporaryDirectory() as tmp:
            os.makedirs(os.path.join(tmp, "Annotations"))
            names = [f"img{i:04d}" for i in range(n_img)]
            with open(os.path.join(tmp, "set.txt"), "w") as f:
                f.write("\n".join(names) + "\n")
            gt = {}
            for nm in names:
                objs = []
                for _ in range(rng.randint(0, 9)):
                    x1, y1 = rng.randint(0, 900), rng.randint(0, 400)
                    w, h = rng.randint(8, 300), rng.randint(8, 200)
                    cls = CLASSES[rng.randint(0, 3)]
                    if cls == empty_class:
                        cls = "car"
                    objs.append((cls, rng.rand() < 0.15, [x1, y1, x1 + w, y1 + h]))
                gt[nm] = objs
                write_xml(os.path.join(tmp, "Annotations", nm + ".xml"), objs)
            per_class = {}
            for cls in CLASSES:
                # detections: jittered ground truth (several per box: duplicates must become false positives) + noise;
                # confidences pairwise distinct AFTER the evaluator's %.3f rounding (np.argsort is unstable on ties)
                dets = []
                for nm in names:
                    for c, _, b in gt[nm]:
                        if c != cls:
                            continue
                        for _ in range(rng.randint(0, 3)):
                            j = rng.randn(4) * 6
                            dets.append((nm, [b[0] + j[0], b[1] + j[1], b[2] + j[2], b[3] + j[3]]))
                while len(dets) < n_det // 3:
                    x1, y1 = rng.rand() * 900, rng.rand() * 400
                    dets.append((names[rng.randint(0, n_img)], [x1, y1, x1 + 8 + rng.rand() * 300, y1 + 8 + rng.rand() * 200]))
                dets = dets[: min(len(dets), 999)]
                conf = (rng.permutation(999)[: len(dets)] + 1) / 1000.0
                lines = [f"{nm} {c:.3f} {b[0] + 1:.1f} {b[1] + 1:.1f} {b[2]:.1f} {b[3]:.1f}" for (nm, b), c in zip(dets, conf)]
                with open(os.path.join(tmp, f"det_{cls}.txt"), "w") as f:
                    f.write("\n".join(lines) + ("\n" if lines else ""))
                out = {}
                for thr, m07 in ((0.5, False), (0.5, True), (0.75, False)):
                    mod.parse_rec.cache_clear()
                    if not lines:
                        continue
                    rec, prec, ap = mod.voc_eval(os.path.join(tmp, "det_{}.txt"), os.path.join(tmp, "Annotations", "{}.xml"),
                                                 os.path.join(tmp, "set.txt"), cls, ovthresh=thr, use_07_metric=m07)
                    out[(thr, m07)] = {"rec": torch.from_numpy(np.asarray(rec)), "prec": torch.from_numpy(np.asarray(prec)),
                                       "ap": float(ap)}
                # what the evaluator parsed back from the text files
                parsed = [ln.split(" ") for ln in lines]
                per_class[cls] = {
                    "det_image": torch.tensor([names.index(p[0]) for p in parsed], dtype=torch.int64),
                    "det_conf": torch.tensor([float(p[1]) for p in parsed], dtype=torch.float64),
                    "det_boxes": torch.tensor([[float(z) for z in p[2:]] for p in parsed], dtype=torch.float64).reshape(-1, 4),
                    "gt_boxes": [torch.tensor([b for c, _, b in gt[nm] if c == cls], dtype=torch.float64).reshape(-1, 4) for nm in names],
                    "gt_difficult": [torch.tensor([bool(d) for c, d, _ in gt[nm] if c == cls], dtype=torch.bool) for nm in names],
                    "out": out}
            cases.append({"id": case_id, "classes": per_class})
    torch.save({"cases": cases, "source": "coin/evaluation/cloud_pascal_voc_evaluation.py:205-319 (voc_eval), 173-202 (voc_ap)"},
               os.path.join(HERE, "voc_eval_ref.pt"))
    print("voc_eval_ref.pt:", sum(len(c["classes"]) for c in cases), "class evaluations")


if __name__ == "__main__":
    main()

"""the VOC evaluator's matching loop. CPU: the oracle's restatement equals the outputs of the
reference's unmodified voc_eval (tests/golden/voc_eval_ref.pt, frozen by tests/golden/make_golden_voc.py). GPU: the device
evaluator equals them too - rec / prec bit for bit (they are ratios of integer counts), AP to 1e-12."""
import math
import warnings

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import coin_ref

GOLD = load_golden("voc_eval_ref.pt")


def _evals():
    for case in GOLD["cases"]:
        for cls, c in case["classes"].items():
            for (thr, m07), want in c["out"].items():
                yield case["id"], cls, c, thr, m07, want


def test_oracle_equals_reference_voc_eval():
    n = 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for cid, cls, c, thr, m07, want in _evals():
            rec, prec, ap = coin_ref.voc_eval_class(c["det_image"].numpy(), c["det_conf"].numpy(), c["det_boxes"].numpy(),
                                                    [b.numpy() for b in c["gt_boxes"]], [d.numpy() for d in c["gt_difficult"]],
                                                    thr, m07)
            assert np.array_equal(rec, want["rec"].numpy(), equal_nan=True), (cid, cls, thr, m07)
            assert np.array_equal(prec, want["prec"].numpy(), equal_nan=True), (cid, cls, thr, m07)
            assert ap == want["ap"] or (math.isnan(ap) and math.isnan(want["ap"]))
            n += 1
    assert n == 27


@pytest.mark.gpu
def test_device_evaluator_equals_reference_voc_eval(dev):
    from coin_b200 import evaluation
    for cid, cls, c, thr, m07, want in _evals():
        gb, go, gd = evaluation.pack_ground_truth(c["gt_boxes"], c["gt_difficult"], dev)
        rec, prec, ap = evaluation.voc_eval_class(c["det_image"].to(dev), c["det_conf"].to(dev), c["det_boxes"].to(dev), gb, go, gd,
                                                  thr, m07)
        assert torch.equal(torch.nan_to_num(rec.cpu(), nan=-1.0), torch.nan_to_num(want["rec"], nan=-1.0)), (cid, cls, thr, m07)
        assert torch.equal(torch.nan_to_num(prec.cpu(), nan=-1.0), torch.nan_to_num(want["prec"], nan=-1.0)), (cid, cls, thr, m07)
        assert (math.isnan(ap) and math.isnan(want["ap"])) or abs(ap - want["ap"]) <= 1e-12, (cid, cls, ap, want["ap"])


@pytest.mark.gpu
def test_device_evaluator_large_and_ties(dev):
    """20 000 detections over 500 images against the oracle (stable order on confidence ties), and argsort_desc alone."""
    from coin_b200 import evaluation
    rng = np.random.RandomState(7)
    n_img, nd = 500, 20000
    gts = [torch.from_numpy(np.concatenate((xy := rng.rand(k, 2) * 800, xy + 10 + rng.rand(k, 2) * 200), 1)).round()
           for k in rng.randint(0, 12, n_img)]
    diff = [torch.from_numpy(rng.rand(len(b)) < 0.1) for b in gts]
    det_image = torch.from_numpy(rng.randint(0, n_img, nd))
    base = torch.stack([gts[i][rng.randint(0, len(gts[i]))] if len(gts[i]) else torch.tensor([5., 5., 50., 50.], dtype=torch.float64)
                        for i in det_image.tolist()])
    det_boxes = base + torch.from_numpy(rng.randn(nd, 4) * 8)
    conf = torch.from_numpy(np.round(rng.rand(nd), 3))           # 3 decimals: thousands of exact ties
    order = evaluation.argsort_desc(conf.to(dev))
    assert torch.equal(order.cpu(), torch.from_numpy(np.argsort(-conf.numpy(), kind="stable")))
    gb, go, gd = evaluation.pack_ground_truth(gts, diff, dev)
    rec, prec, ap = evaluation.voc_eval_class(det_image.to(dev), conf.to(dev), det_boxes.to(dev), gb, go, gd, 0.5, False)
    wr, wp, wap = coin_ref.voc_eval_class(det_image.numpy(), conf.numpy(), det_boxes.numpy(), [b.numpy() for b in gts],
                                          [d.numpy() for d in diff], 0.5, False)
    assert np.array_equal(rec.cpu().numpy(), wr) and np.array_equal(prec.cpu().numpy(), wp)
    assert abs(ap - wap) <= 1e-12

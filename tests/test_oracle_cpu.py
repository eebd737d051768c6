"""CPU suite (-m "not gpu"): pins the oracle against the golden vectors and the torchvision CPU
operators of this image, checks the host logic, and that the C-ABI library exports what
include/coinops.h declares. No CUDA kernel runs here."""
import ctypes
import os
import re

import pytest
import torch
import torchvision

from conftest import ROOT, load_golden
from coin_b200 import synth
from oracle import clib, coin_ref, d2_ref


# ---------------------------------------------------------------------------------------------
# C restatement vs torchvision CPU ops (fresh) and vs the frozen vectors
# ---------------------------------------------------------------------------------------------
def test_c_roi_align_matches_golden_bitwise():
    tv = load_golden("tv_ops.pt")
    for case in tv["roi_align"]:
        out = clib.roi_align_fwd(tv["x"], tv["rois"], 1.0 / 16, case["ph"], case["pw"], case["sr"], case["aligned"])
        assert torch.equal(out, case["out"]), case
        gin = clib.roi_align_bwd(case["grad_out"], tv["rois"], 1.0 / 16, case["ph"], case["pw"], *tv["x"].shape,
                                 case["sr"], case["aligned"])
        torch.testing.assert_close(gin, case["grad_in"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("pooled,aligned,sr", [(7, True, 0), (14, True, 0), (7, False, 2)])
def test_c_roi_align_matches_torchvision_foggy_shape(pooled, aligned, sr):
    g = synth.gen(11)
    shape = synth.SHAPES["foggy_cpu"]
    h, w = shape.feat_hw
    x = torch.randn(2, 16, h, w, generator=g)
    boxes = synth.random_boxes(g, 64, shape.height, shape.width)
    rois = torch.cat((torch.randint(0, 2, (64, 1), generator=g).float(), boxes), dim=1)
    ref = torchvision.ops.roi_align(x, rois, (pooled, pooled), 1.0 / 16, sr, aligned)
    out = clib.roi_align_fwd(x, rois, 1.0 / 16, pooled, pooled, sr, aligned)
    assert torch.equal(out, ref)


def test_c_nms_and_iou_match_torchvision():
    tv = load_golden("tv_ops.pt")["nms"]
    for thr in (0.5, 0.7):
        keep = clib.nms(tv["boxes"], tv["scores"], thr)
        assert torch.equal(keep, tv[f"keep_{thr}"])
        assert torch.equal(keep, torchvision.ops.nms(tv["boxes"], tv["scores"], thr))
    iou = clib.pairwise_iou(tv["boxes"][:40], tv["boxes"][40:110])
    assert torch.equal(iou, tv["iou"])
    assert torch.equal(iou, d2_ref.pairwise_iou(tv["boxes"][:40], tv["boxes"][40:110]))


def test_pairwise_iou_restatement_is_bitwise_box_iou_and_zero_on_degenerate():
    g = synth.gen(3)
    a = synth.random_boxes(g, 100, 600, 1200)
    b = d2_ref.grid_anchors(37, 75, 16, d2_ref.cell_anchors())
    assert b.shape[0] == 41625
    assert torch.equal(d2_ref.pairwise_iou(a, b), torchvision.ops.box_iou(a, b))
    z = torch.tensor([[5.0, 5.0, 5.0, 5.0]])
    assert d2_ref.pairwise_iou(z, z).item() == 0.0  # torchvision gives NaN here; detectron2 gives 0


def test_batched_nms_golden_and_strategy_switch():
    tv = load_golden("tv_ops.pt")
    s, b = tv["nms"], tv["nms_big"]
    assert d2_ref.batched_nms_strategy(300) == "trick" and d2_ref.batched_nms_strategy(1500) == "vanilla"
    assert torch.equal(d2_ref.batched_nms(s["boxes"], s["scores"], s["idxs"], 0.5), s["batched_keep_0.5"])
    assert torch.equal(d2_ref.batched_nms(b["boxes"], b["scores"], b["idxs"], 0.5), b["batched_keep_0.5"])


def test_nms_threshold_is_strict_and_sort_is_stable():
    boxes = torch.tensor([[0.0, 0.0, 10.0, 10.0], [0.0, 0.0, 10.0, 5.0], [0.0, 0.0, 10.0, 10.0], [50.0, 50.0, 60.0, 60.0]])
    scores = torch.tensor([0.9, 0.9, 0.9, 0.9])
    assert d2_ref.pairwise_iou(boxes[:1], boxes[1:2]).item() == 0.5
    assert clib.nms(boxes, scores, 0.5).tolist() == [0, 1, 3]  # IoU == thr is kept; ties -> lower index first
    assert d2_ref.nms(boxes, scores, 0.5).tolist() == [0, 1, 3]


# ---------------------------------------------------------------------------------------------
# restated COIN code vs the reference's own coin/layers/nms.py (frozen outputs)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("method", ["ms", "ma", "ps", "pa", "pm", "as", "aa", "am", "nms", "mm"])
def test_mynms_restatement_matches_reference_outputs(method):
    gold = load_golden(f"fusion_nms_{method}.pt")
    for case in gold["cases"]:
        keep, ob, os_, op, ol = coin_ref.mynms(method, case["boxes"], case["scores"], case["probs"], case["labels"],
                                               case["thr"])
        assert torch.equal(keep, case["keep"])
        assert torch.equal(ol, case["out_classes"])
        assert torch.equal(ob, case["out_boxes"])
        assert torch.equal(os_, case["out_scores"])
        assert torch.equal(op, case["out_probs"])


def test_mynms_invariants_from_reference_asserts():
    # gdino_processor.py:284-285: after 'ms' fusion score == max prob and class == argmax prob
    case = load_golden("fusion_nms_ms.pt")["cases"][0]
    assert torch.equal(case["out_scores"], case["out_probs"].max(1)[0])
    assert torch.equal(case["out_classes"], case["out_probs"].max(1)[1])
    assert bool((case["out_scores"][:-1] >= case["out_scores"][1:]).all())


# ---------------------------------------------------------------------------------------------
# detectron2 restatements: hand-computed cases and properties
# ---------------------------------------------------------------------------------------------
def test_matcher_cases():
    q = torch.tensor([[0.1, 0.6, 0.4, 0.0], [0.1, 0.2, 0.8, 0.0], [0.05, 0.6, 0.3, 0.0]])
    m = d2_ref.Matcher([0.5], [0, 1], False)
    idx, lab = m(q)
    assert idx.tolist() == [0, 0, 1, 0] and lab.tolist() == [0, 1, 1, 0]  # ties -> first row
    m = d2_ref.Matcher([0.3, 0.7], [0, -1, 1], True)
    idx, lab = m(q)
    assert idx.tolist() == [0, 0, 1, 0]
    assert lab.tolist() == [0, 1, 1, 0]  # col1 low-quality (row0 & row2 max 0.6), col2 is > 0.7
    empty_idx, empty_lab = m(torch.zeros(0, 5))
    assert empty_idx.tolist() == [0] * 5 and empty_lab.tolist() == [0] * 5
    # a GT row whose best IoU is 0 marks every zero column positive (documented d2 behaviour)
    q0 = torch.tensor([[0.0, 0.0, 0.0], [0.2, 0.0, 0.9]])
    assert m(q0)[1].tolist() == [1, 1, 1]


def test_box2box_roundtrip_and_clamp():
    g = synth.gen(5)
    src = synth.random_boxes(g, 200, 600, 1200)
    tgt = synth.jitter(g, src, 0.2, 600, 1200)
    t = d2_ref.Box2BoxTransform((10.0, 10.0, 5.0, 5.0))
    torch.testing.assert_close(t.apply_deltas(t.get_deltas(src, tgt), src), tgt, rtol=1e-4, atol=1e-3)
    big = t.apply_deltas(torch.tensor([[0.0, 0.0, 100.0, 100.0]]), torch.tensor([[0.0, 0.0, 16.0, 16.0]]))
    assert abs((big[0, 2] - big[0, 0]).item() - 1000.0) < 1e-2  # exp(log(1000/16)) * 16


def test_pooler_levels_and_single_level_equivalence():
    boxes = torch.tensor([[0.0, 0.0, 224.0, 224.0], [0.0, 0.0, 112.0, 112.0], [0.0, 0.0, 448.0, 448.0],
                          [0.0, 0.0, 10.0, 10.0], [0.0, 0.0, 2000.0, 2000.0], [0.0, 0.0, 223.9, 223.9]])
    assert d2_ref.assign_boxes_to_levels([boxes], 2, 5).tolist() == [2, 1, 3, 0, 3, 1]
    g = synth.gen(9)
    x = torch.randn(2, 4, 20, 30, generator=g)
    bl = [synth.random_boxes(g, 5, 320, 480), synth.random_boxes(g, 3, 320, 480)]
    out = d2_ref.roi_pooler([x], bl, 7, (1.0 / 16,))
    rois = d2_ref.pooler_format(bl)
    assert torch.equal(out, torchvision.ops.roi_align(x, rois, (7, 7), 1.0 / 16, 0, True))


# ---------------------------------------------------------------------------------------------
# knowledge separation (restated trainer.py:338-485): hand cases, invariants, policies
# ---------------------------------------------------------------------------------------------
def _dets(boxes, classes, scores, k1=4):
    boxes = torch.tensor(boxes, dtype=torch.float32).reshape(-1, 4)
    classes = torch.tensor(classes, dtype=torch.int64)
    scores = torch.tensor(scores, dtype=torch.float32)
    probs = torch.zeros(len(classes), k1)
    if len(classes):
        probs[torch.arange(len(classes)), classes] = scores
    return {"gt_boxes": boxes, "gt_classes": classes, "scores": scores, "probs": probs}


def test_match_dual_teacher_hand_case():
    online = _dets([[0, 0, 10, 10], [100, 100, 120, 120], [200, 200, 230, 230]], [1, 2, 0], [0.9, 0.8, 0.7])
    offline = _dets([[1, 1, 10, 10], [101, 101, 121, 121], [300, 300, 320, 320]], [1, 0, 2], [0.6, 0.5, 0.4])
    a, b, c = coin_ref.match_dual_teacher(online, offline, "RCNN")
    assert a["gt_boxes"].tolist() == [[0, 0, 10, 10]] and a["gt_classes"].tolist() == [1]
    assert b["gt_boxes"].tolist() == [[100, 100, 120, 120]]
    assert b["gt_classes_online"].tolist() == [2] and b["gt_classes_offline"].tolist() == [0]
    assert c["gt_boxes"].tolist() == [[300, 300, 320, 320], [200, 200, 230, 230]]  # offline-only first
    a2, b2, c2 = coin_ref.match_dual_teacher(online, offline, "RPN")
    assert b2 is None and len(a2["gt_boxes"]) == 2 and torch.equal(c2["gt_boxes"], c["gt_boxes"])
    a3, _, _ = coin_ref.match_dual_teacher(online, offline, "RCNN", weight_for_box_a=0.5)
    w = torch.tensor([0.9, 0.6]) / 1.5
    torch.testing.assert_close(a3["gt_boxes"][0], online["gt_boxes"][0] * w[0] + offline["gt_boxes"][0] * w[1])


def test_match_dual_teacher_empty_sides():
    online = _dets([[0, 0, 10, 10]], [1], [0.9])
    offline = _dets([[0, 0, 10, 10], [50, 50, 70, 70]], [1, 2], [0.95, 0.3])
    empty = _dets([], [], [])
    a, b, c = coin_ref.match_dual_teacher(empty, offline, "RCNN")
    assert a["gt_boxes"].tolist() == [[0, 0, 10, 10]] and len(b["gt_boxes"]) == 0
    assert c["gt_boxes"].tolist() == [[50, 50, 70, 70]]
    a, b, c = coin_ref.match_dual_teacher(online, empty, "RCNN")
    assert a["gt_boxes"].tolist() == [[0, 0, 10, 10]] and len(c["gt_boxes"]) == 0
    a, b, c = coin_ref.match_dual_teacher(empty, empty, "RPN")
    assert len(a["gt_boxes"]) == 0 and b is None and len(c["gt_boxes"]) == 0


def test_match_dual_teacher_duplicates_and_policies_on_synthetic():
    shape = synth.SHAPES["foggy_cpu"]
    batch = synth.image_batch(shape)
    for img in batch["images"]:
        cloud, clip = img["cloud"], img["clip"]
        uniq, groups = coin_ref.delete_duplicate_boxes(clip, return_split=True)
        assert len(groups) > 0 and coin_ref.length(uniq) + sum(coin_ref.length(x) for x in groups) == shape.clip
        assert len(coin_ref.self_clusters(cloud["gt_boxes"], 0.95)) > 0
        for tag in ("RCNN", "RPN"):
            a, b, c = coin_ref.match_dual_teacher(cloud, clip, tag)
            # every cloud box ends in exactly one of: a common pair or the private set
            assert len(a["gt_boxes"]) > 0 and len(c["gt_boxes"]) > 0
            if tag == "RCNN":
                assert bool((b["gt_classes_online"] != b["gt_classes_offline"]).all())
                clash = torch.eq(b["gt_boxes"].unsqueeze(1), a["gt_boxes"]).sum(-1) == 4
                assert not bool(clash.any())
            # The literal CPython-set policy yields the same private SET (row order may differ). A/B can
            # legitimately differ: util.py:497 keys its decision on the FIRST box of a self-cluster, and
            # "first" is CPython hash order in the reference (e.g. list({7, 9}) == [9, 7]).
            a2, b2, c2 = coin_ref.match_dual_teacher(cloud, clip, tag, set_order="cpython")
            assert sorted(map(tuple, c["gt_boxes"].tolist())) == sorted(map(tuple, c2["gt_boxes"].tolist()))
            assert abs(len(a["gt_boxes"]) - len(a2["gt_boxes"])) <= len(coin_ref.self_clusters(cloud["gt_boxes"], 0.95))


def test_fast_rcnn_inference_restatement_properties():
    g = synth.gen(21)
    r, k1 = 300, 9
    rois = synth.random_boxes(g, r, 600, 1200)
    deltas, logits = synth.deltas_scores(g, r, k1)
    boxes = d2_ref.Box2BoxTransform((10.0, 10.0, 5.0, 5.0)).apply_deltas(deltas, rois)
    probs = torch.softmax(logits, dim=1)
    probs[5, 2] = float("nan")
    res, kept = coin_ref.fast_rcnn_inference_single_image(boxes, probs, (600, 1200), 0.05, 0.5, 100)
    assert len(res["scores"]) <= 100 and bool((res["scores"][:-1] >= res["scores"][1:]).all())
    assert bool((res["scores"] > 0.05).all()) and int(kept.max()) < r - 1  # indices are in the filtered frame
    assert bool((res["pred_boxes"][:, 0] >= 0).all()) and bool((res["pred_boxes"][:, 2] <= 1200).all())


def test_relabel_epilogues():
    idx = torch.tensor([0, 1, 2, 3, 4, 4])
    lab = torch.tensor([1, 0, 1, 1, 0, 1], dtype=torch.int8)
    assert coin_ref.relabel_roi(idx, lab, 2, 1, 2).tolist() == [1, 0, 1, -1, 0, -1]
    l, i, di, dl = coin_ref.relabel_rpn(idx, lab, 3, 2)
    assert l.tolist() == [1, 0, 1, -1, 0, -1] and i.tolist() == [0, 1, 2, 0, 0, 0]
    assert di.tolist() == [0, 0, 0, 0, 0, 1] and dl.tolist() == [0, 0, 0, 1, 0, 1]


def test_process_boxes_scale_flip():
    b = torch.tensor([[100.0, 50.0, 300.0, 250.0]])
    out = coin_ref.process_boxes(b, (1024, 2048), (600, 1200), "horizontal")
    sx, sy = 1200 / 2048, 600 / 1024
    torch.testing.assert_close(out, torch.tensor([[1200 - 300 * sx, 50 * sy, 1200 - 100 * sx, 250 * sy]]))


# ---------------------------------------------------------------------------------------------
def test_predict_proposals_single_properties():
    """d2 RPN.predict_proposals restatement (one image, one level): a hand case, and on random input the invariants
    of find_top_rpn_proposals - descending logits, boxes inside the image and larger than min_box_size, pairwise
    IoU <= nms_thresh, at most post_nms_topk rows, non-finite rows dropped (FloatingPointError when training)."""
    anchors = torch.tensor([[0.0, 0.0, 10.0, 10.0], [0.0, 0.0, 10.0, 10.0], [20.0, 20.0, 40.0, 40.0], [100.0, 100.0, 130.0, 130.0]])
    deltas = torch.zeros(4, 4)
    deltas[1, 0] = 0.05                      # shifted by half a pixel: IoU with box 0 > 0.7 -> suppressed
    logits = torch.tensor([2.0, 1.0, 3.0, 0.5])
    boxes, sc = d2_ref.predict_proposals_single(anchors, deltas, logits, (50, 50), 0.7, 4, 10)
    assert sc.tolist() == [3.0, 2.0]         # box 3 lies outside the 50x50 image: clipped to empty
    assert torch.equal(boxes, torch.tensor([[20.0, 20.0, 40.0, 40.0], [0.0, 0.0, 10.0, 10.0]]))
    g = torch.Generator().manual_seed(11)
    base = d2_ref.grid_anchors(10, 12, 16, d2_ref.cell_anchors((32, 64), (0.5, 1.0, 2.0)))
    d = 0.3 * torch.randn(base.shape[0], 4, generator=g)
    lg = torch.randn(base.shape[0], generator=g)
    lg[5] = float("nan")
    d[9, 2] = float("inf")
    boxes, sc = d2_ref.predict_proposals_single(base, d, lg, (160, 192), 0.7, 300, 50, min_box_size=2.0)
    assert len(boxes) <= 50 and torch.isfinite(boxes).all() and torch.isfinite(sc).all()
    assert (sc[:-1] >= sc[1:]).all()
    assert (boxes[:, 0] >= 0).all() and (boxes[:, 2] <= 192).all() and (boxes[:, 3] <= 160).all()
    assert ((boxes[:, 2] - boxes[:, 0]) > 2.0).all() and ((boxes[:, 3] - boxes[:, 1]) > 2.0).all()
    iou = torchvision.ops.box_iou(boxes, boxes)
    iou.fill_diagonal_(0)
    assert float(iou.max()) <= 0.7
    with pytest.raises(FloatingPointError):
        d2_ref.predict_proposals_single(base, d, lg, (160, 192), 0.7, 300, 50, training=True)


def test_detector_postprocess_and_box_reg_loss_hand_cases():
    """Hand-computed cases for the restatements of detectron2's detector_postprocess (<- clip_rcnn.py:424) and of
    FastRCNNOutputLayers.box_reg_loss (fast_rcnn.py:601-646)."""
    boxes = torch.tensor([[10.0, 20.0, 110.0, 220.0], [590.0, 10.0, 700.0, 50.0], [5.0, 5.0, 5.0, 9.0]])
    out, keep = d2_ref.detector_postprocess(boxes, (300, 600), 600, 1200)      # x2 in both directions
    assert keep.tolist() == [True, True, False]                                 # zero-width box dropped
    assert torch.equal(out, torch.tensor([[20.0, 40.0, 220.0, 440.0], [1180.0, 20.0, 1200.0, 100.0]]))  # clipped to w=1200
    # one foreground proposal identical to its GT except a shift of dx = w/10: target delta = (wx * 0.1, 0, 0, 0)
    props = torch.tensor([[0.0, 0.0, 100.0, 50.0], [0.0, 0.0, 10.0, 10.0]])
    gts = torch.tensor([[10.0, 0.0, 110.0, 50.0], [0.0, 0.0, 10.0, 10.0]])
    cls = torch.tensor([2, 8])                                                   # second row is background (K = 8)
    pred = torch.zeros(2, 4)
    loss = coin_ref.box_reg_loss((10.0, 10.0, 5.0, 5.0), props, gts, pred, cls, 8)
    assert abs(float(loss) - 1.0 / 2) < 1e-6                                     # |0 - 10*0.1| summed, / R = 2 regions
    loss = coin_ref.box_reg_loss((10.0, 10.0, 5.0, 5.0), props, gts, pred, cls, 8, smooth_l1_beta=2.0, normalizer=4.0)
    assert abs(float(loss) - 0.5 * 1.0 / 2.0 / 4.0) < 1e-6                       # quadratic branch: 0.5 * n^2 / beta
    per_class = torch.zeros(2, 32)
    per_class[0, 8:12] = torch.tensor([1.0, 0.0, 0.0, 0.0])                      # class 2's slot already equals the target
    assert float(coin_ref.box_reg_loss((10.0, 10.0, 5.0, 5.0), props, gts, per_class, cls, 8)) < 1e-6


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's CPU path = the oracle port, timed on the host cores) prints ONE JSON
    line with the keys the driver reads, and does not need a GPU."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "roi_path_images_per_sec" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"] == "foggy_roi_head"


# the C ABI: the library loads and exports every symbol include/coinops.h declares
# ---------------------------------------------------------------------------------------------
def test_abi_exports_match_header():
    header = open(os.path.join(ROOT, "include", "coinops.h")).read()
    declared = set(re.findall(r"\b(coin_[a-z0-9_]+)\s*\(", header))
    declared -= {"coin_stream_t", "coin_level_t"}
    from coin_b200 import _lib
    lib = ctypes.CDLL(_lib.SO_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in coinops.h but not exported"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert lib.coin_version() >= 100


def test_no_cpu_fallback():
    import coin_b200
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        coin_b200.pairwise_iou(coin_b200.Boxes(torch.zeros(2, 4)), coin_b200.Boxes(torch.zeros(2, 4)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        coin_b200.ROIAlign(7, 1.0 / 16, 0)(torch.zeros(1, 4, 8, 8), torch.zeros(1, 5))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "coin_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f"{f} imports the oracle"

"""Runs each latency-bound stage of the step ONCE, eagerly (for `ncu --metrics gpu__time_duration.sum`: the launch list
with per-kernel durations of det_postprocess, the RPN NMS, match_abc and the two labelling stages)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coin_b200 import ops, pipeline, synth  # noqa: E402

dev = torch.device("cuda:0")
shape = synth.SHAPES[sys.argv[1] if len(sys.argv) > 1 else "foggy_roi_head"]
step = pipeline.RoIPathStep(shape, dev)
d = step.to_device(synth.image_batch(shape))
img = (shape.height, shape.width)
for rep in range(2):
    torch.cuda.nvtx.range_push(f"rep{rep}")
    dec = ops.apply_deltas(d["0.teacher_deltas"], d["0.teacher_rois"], step.BBOX_WEIGHTS, clip_to=img)
    b, s_, p, c, roi, nd = ops.det_postprocess(dec, d["0.teacher_probs"], img, 0.05, 0.5, 100, sync=False)
    ops.batched_nms(d["0.rpn_boxes"], d["0.rpn_scores"], None, 0.7, "plain", shape.rpn_post_nms, sync=False)
    cloud = {"gt_boxes": d["0.cloud.gt_boxes"] / pipeline.ORIG_SCALE, "gt_classes": d["0.cloud.gt_classes"],
             "scores": d["0.cloud.scores"], "probs": d["0.cloud.probs"]}
    clip = {"gt_boxes": b, "gt_classes": c, "scores": s_, "probs": p}
    both = ops.match_abc_fields_both_dev(cloud, clip, nd, 0.5, 1.0)
    a, bb, cc, cnt = both["RCNN"]
    n_a, n_b, n_c = cnt[0:1], cnt[1:2], cnt[2:3]
    gt, n_gt = ops.concat_rows([(a["gt_boxes"], n_a, 0.0), (bb["gt_boxes"], n_b, 0.0), (cc["gt_boxes"], n_c, 0.0)])
    idx2, lab2 = ops.iou_match_dev(gt, n_gt, step.anchors, None, [0.3, 0.7], [0, -1, 1], True)
    ops.relabel_rpn_dev_(idx2, lab2, n_a, n_c)
    props, n_props = ops.concat_rows([(d["0.proposals"], None, 0.0), (a["gt_boxes"], n_a, 0.0), (bb["gt_boxes"], n_b, 0.0)])
    idx, lab = ops.iou_match_dev(gt, n_gt, props, n_props, [0.5], [0, 1], False)
    ops.relabel_roi_dev_(idx, lab, n_props, n_a, n_b, n_c)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
print("counts", cnt.tolist(), "ndet", int(nd))

"""Times the latency-bound stages of the step alone (CUDA events, 30 iterations each) on the bench shape."""
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


def timeit(name, fn, iters=30):
    """Device time of the op's launch chain: captured once in a CUDA graph (no host launch overhead), replayed."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        keep = fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    print(f"{s.elapsed_time(e) / iters * 1e3:9.1f} us  {name}", flush=True)
    del keep


dec = ops.apply_deltas(d["0.teacher_deltas"], d["0.teacher_rois"], step.BBOX_WEIGHTS, clip_to=img)
timeit("det_postprocess (1000 RoIs x 8 classes, sync-free)",
       lambda: ops.det_postprocess(dec, d["0.teacher_probs"], img, 0.05, 0.5, 100, sync=False))
b, s_, p, c, roi, nd = ops.det_postprocess(dec, d["0.teacher_probs"], img, 0.05, 0.5, 100, sync=False)
timeit(f"batched_nms plain {shape.rpn_pre_nms} boxes thr 0.7 max_keep {shape.rpn_post_nms}",
       lambda: ops.batched_nms(d["0.rpn_boxes"], d["0.rpn_scores"], None, 0.7, "plain", shape.rpn_post_nms, sync=False))
cloud = {"gt_boxes": d["0.cloud.gt_boxes"] / pipeline.ORIG_SCALE, "gt_classes": d["0.cloud.gt_classes"],
         "scores": d["0.cloud.scores"], "probs": d["0.cloud.probs"]}
clip = {"gt_boxes": b, "gt_classes": c, "scores": s_, "probs": p}
for tag in ("RCNN", "RPN"):
    timeit(f"match_abc_fields_dev tag {tag}", lambda: ops.match_abc_fields_dev(cloud, clip, nd, tag, 0.5, 1.0))
timeit("match_abc_fields_both_dev (both tags: 1 match launch + 2 pack launches)",
       lambda: ops.match_abc_fields_both_dev(cloud, clip, nd, 0.5, 1.0))
a, bb, cc, cnt = ops.match_abc_fields_dev(cloud, clip, nd, "RCNN", 0.5, 1.0)
n_a, n_b, n_c = cnt[0:1], cnt[1:2], cnt[2:3]
gt, n_gt = ops.concat_rows([(a["gt_boxes"], n_a, 0.0), (bb["gt_boxes"], n_b, 0.0), (cc["gt_boxes"], n_c, 0.0)])
timeit("concat_rows (3 segments)", lambda: ops.concat_rows([(a["gt_boxes"], n_a, 0.0), (bb["gt_boxes"], n_b, 0.0), (cc["gt_boxes"], n_c, 0.0)]))
timeit("iou_match_dev gt x 41625 anchors, low quality", lambda: ops.iou_match_dev(gt, n_gt, step.anchors, None, [0.3, 0.7], [0, -1, 1], True))
props, n_props = ops.concat_rows([(d["0.proposals"], None, 0.0), (a["gt_boxes"], n_a, 0.0), (bb["gt_boxes"], n_b, 0.0)])
timeit("iou_match_dev gt x proposals", lambda: ops.iou_match_dev(gt, n_gt, props, n_props, [0.5], [0, 1], False))
print("counts", cnt.tolist(), "ndet", int(nd))

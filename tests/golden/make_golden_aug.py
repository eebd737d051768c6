"""Freezes GDINO_PROCESSOR.post_process with an 'AUG' detection set (gdino_processor.py:287-298; run in the BUILD container only):

    python tests/golden/make_golden_aug.py      ->  tests/golden/gdino_aug_ref.pt

The AUG branch (detections of an augmented view, cfg.INPUT.TEACHER_CLOUD.COLLECT_AUG; off in the paper's runs) appends the AUG
set to the NMS'ed RPN set and runs `mynms.nms` over the union: the 'RPN_AUG' tag that BASE_Trainer.preprocess_results
(base.py:128-136) then prefers over 'RPN'. Executed unmodified through oracle/ref_loader.py like make_golden_ref.py."""
import os
import types

import torch

import make_golden_ref as G
from make_golden_ref import Instances, Recorder, gproc, synth, to_dict, to_inst

HERE = os.path.dirname(os.path.abspath(__file__))


def dets(g, n, objs):
    boxes = synth.jitter(g, objs[torch.randint(0, len(objs), (n,), generator=g)], 0.05, 1024, 2048)
    logits = 2.0 * torch.randn(n, 9, generator=g)
    logits[:, -1] = -float("inf")
    probs = torch.softmax(logits, 1)
    return {"pred_boxes": boxes, "scores": probs.max(1)[0], "pred_classes": probs.argmax(1), "probs": probs}


def gen():
    cases = []
    for method in ("ps", "nms"):
        gproc.mynms.update(method)
        for seed, n, n_aug in ((51, 120, 60), (52, 9, 0), (53, 0, 7)):
            g = synth.gen(seed)
            objs = synth.random_boxes(g, 12, 1024, 2048)
            ori, aug = dets(g, n, objs), dets(g, n_aug, objs)
            me = Recorder(RCNN_THRESH=0.45, RPN_THRESH=0.3, COLLECT_NMS_THRESH=0.6)
            me.nms = types.MethodType(gproc.GDINO_PROCESSOR.nms, me)
            me.draw = lambda *a, **k: None
            me.save_path = "/tmp"
            outputs = {"ORI": {"instances": to_inst(ori, (1024, 2048), cls=Instances)},
                       "AUG": {"instances": to_inst(aug, (1024, 2048), cls=Instances)}}
            gproc.GDINO_PROCESSOR.post_process(me, outputs, [{"file_name": "x/y.png"}])
            cases.append({"method": method, "in": ori, "aug": aug, "rcnn_thresh": 0.45, "rpn_thresh": 0.3, "nms_thresh": 0.6,
                          "out": {k: to_dict(outputs[k]["instances"]) for k in ("RCNN", "RPN", "RPN_AUG")}})
    gproc.mynms.update("nms")
    torch.save({"cases": cases, "source": "coin/modeling/meta_arch/gdino_processor.py:164-182,287-298"},
               os.path.join(HERE, "gdino_aug_ref.pt"))
    return len(cases)


if __name__ == "__main__":
    print("gdino AUG cases:", gen())

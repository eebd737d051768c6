"""Freezes outputs of the reference's 'pre_train' labelling branches (run in the BUILD container only):

    python tests/golden/make_golden_pretrain.py      ->  tests/golden/labels_pretrain_ref.pt

  roi   OpenVocabularyRes5ROIHeads.label_and_sample_proposals(..., 'pre_train')   clip_roi_heads.py:286-340
  rpn   DualTeacherRPN.label_and_sample_anchors(..., 'pre_train')                 rpn.py:139-197

Both are executed unmodified through oracle/ref_loader.py (see make_golden_ref.py) with torch.manual_seed(2024). Each
case exists without and with `no_thresh_boxes` (boxes that must become neither foreground nor background; their producer,
base.py:121, is commented out in the reference, so the stand-in attaches the field directly - it may differ in length
from the other fields, which Instances.set would refuse).
"""
import os
import zlib

import torch

import make_golden_ref as G          # installs the stand-in and loads the reference's modules
from make_golden_ref import Boxes, Instances, Matcher, Recorder, heads, rpn_mod, synth, to_dict

HERE = os.path.dirname(os.path.abspath(__file__))


def target(gt_boxes, classes, probs, size, no_thresh):
    t = Instances(size)
    t.gt_boxes = Boxes(gt_boxes.clone())
    t.gt_classes_offline = classes.clone()
    t.gt_probs = probs.clone()
    if no_thresh is not None:
        t._fields["no_thresh_boxes"] = Boxes(no_thresh.clone())     # (length differs from len(t): see the docstring)
    return t


def gen():
    from detectron2.modeling.roi_heads import ROIHeads
    from detectron2.modeling.proposal_generator import RPN
    from oracle import d2_ref
    out = {"roi": [], "rpn": []}
    for label, shape_name, n_gt, n_nt, n_prop in (("foggy", "foggy_cpu", 30, 12, 2000), ("tiny", "tiny", 6, 3, 200),
                                                  ("no_gt", "tiny", 0, 4, 200), ("no_nt", "tiny", 5, 0, 200)):
        shape = synth.SHAPES[shape_name]
        size = (shape.height, shape.width)
        g = synth.gen(zlib.crc32(("pretrain." + label).encode()) % 1000)
        gt = synth.random_boxes(g, n_gt, shape.height, shape.width)
        nt = synth.random_boxes(g, n_nt, shape.height, shape.width)
        cls = torch.randint(0, 8, (n_gt,), generator=g)
        probs = torch.softmax(torch.randn(n_gt, 9, generator=g), 1)
        objs = torch.cat((gt, nt)) if n_gt + n_nt else synth.random_boxes(g, 4, shape.height, shape.width)
        props = synth.rois_for(g, shape, objs, n_prop)
        logits = torch.randn(n_prop, generator=g)
        hf, wf = shape.feat_hw
        anchors = d2_ref.grid_anchors(hf, wf, 16, d2_ref.cell_anchors())
        for with_nt in (False, True):
            # --- RoI head, Matcher([0.5],[0,1],False), PROPOSAL_APPEND_GT
            me = Recorder(proposal_append_gt=True, proposal_matcher=Matcher([0.5], [0, 1], allow_low_quality_matches=False),
                          num_classes=8, batch_size_per_image=512, positive_fraction=0.25, BG_TRAIN=True)

            def sample(matched_idxs, matched_labels, gt_classes, me=me):
                me.sample_inputs.append((matched_idxs.clone(), matched_labels.clone(), gt_classes.clone()))
                return ROIHeads._sample_proposals(me, matched_idxs, matched_labels, gt_classes)
            me._sample_proposals = sample
            p = Instances(size)
            p.proposal_boxes = Boxes(props.clone())
            p.objectness_logits = logits.clone()
            torch.manual_seed(2024)
            fn = heads.OpenVocabularyRes5ROIHeads.label_and_sample_proposals
            fn = getattr(fn, "__wrapped__", fn)
            res = fn(me, [p], [target(gt, cls, probs, size, nt if with_nt else None)], "pre_train")
            mi, ml, gc = me.sample_inputs[0]
            fg, bg = res[0]
            out["roi"].append({"label": label, "with_no_thresh": with_nt, "gt_boxes": gt, "gt_classes_offline": cls,
                               "gt_probs": probs, "no_thresh_boxes": nt, "proposals": props, "objectness_logits": logits,
                               "image_size": size, "matched_idxs": mi, "matched_labels": ml,
                               "sampled": {"fg": to_dict(fg), "bg": to_dict(bg)}, "torch_seed": 2024, "num_classes": 8,
                               "batch_size_per_image": 512, "positive_fraction": 0.25})
            # --- RPN, Matcher([0.3,0.7],[0,-1,1],True), boundary thresh -1
            me2 = Recorder(anchor_matcher=Matcher([0.3, 0.7], [0, -1, 1], allow_low_quality_matches=True),
                           anchor_boundary_thresh=-1, batch_size_per_image=256, positive_fraction=0.5)

            def subsample(label_vec, me2=me2):
                me2.sample_inputs.append(label_vec.clone())
                return RPN._subsample_labels(me2, label_vec)
            me2._subsample_labels = subsample
            torch.manual_seed(2024)
            gl, gb = rpn_mod.DualTeacherRPN.label_and_sample_anchors(
                me2, [Boxes(anchors.clone())], [target(gt, cls, probs, size, nt if with_nt else None)], "pre_train")
            out["rpn"].append({"label": label, "with_no_thresh": with_nt, "gt_boxes": gt, "no_thresh_boxes": nt,
                               "anchors_hw": (hf, wf), "labels_before_sampling": me2.sample_inputs[0], "gt_labels": gl[0],
                               "matched_gt_boxes_colsum": gb[0].double().sum(0), "matched_gt_boxes_head": gb[0][:256].clone(),
                               "torch_seed": 2024, "batch_size_per_image": 256, "positive_fraction": 0.5})
    torch.save({**out, "source": "coin/modeling/roi_heads/clip_roi_heads.py:286-340, coin/modeling/proposal_generator/rpn.py:"
                                 "139-197"}, os.path.join(HERE, "labels_pretrain_ref.pt"))
    return len(out["roi"]), len(out["rpn"])


if __name__ == "__main__":
    print("pre_train label cases (roi, rpn):", gen())

"""SURVEY 8(f) rank 3 on the device: the collected cloud detections live in the flat arrays of a DetectionCache in HBM;
the step's cloud inputs are views into them (no host copy per lookup / per step) and the step computes what it computes
from host-fed inputs."""
import pytest
import torch

from coin_b200 import integration, pipeline, synth
from coin_b200.cache import DetectionCache
from coin_b200.structures import Boxes, Instances
from oracle import coin_ref

pytestmark = pytest.mark.gpu


def _per_file(batch, shape):
    out, names = {}, []
    for i, img in enumerate(batch["images"]):
        c = img["cloud"]
        inst = Instances((int(shape.height * pipeline.ORIG_SCALE), int(shape.width * pipeline.ORIG_SCALE)))
        inst.pred_boxes = Boxes(c["gt_boxes"] * pipeline.ORIG_SCALE)
        inst.scores, inst.pred_classes, inst.probs = c["scores"], c["gt_classes"], c["probs"]
        name = f"foggy/img{i}.png"
        names.append(name)
        out[name] = {"file_name": name, "image_id": i, "height": inst.image_size[0], "width": inst.image_size[1],
                     "RCNN": {"instances": inst}}
    return out, names


def test_step_reads_cloud_detections_from_the_device_cache(dev, tmp_path):
    shape = synth.SHAPES["tiny"]
    batch = synth.image_batch(shape)
    per_file, names = _per_file(batch, shape)
    # through the cache's own tensor-only file, straight onto the device
    path = str(tmp_path / "cache.pt")
    DetectionCache.from_reference_dict(per_file).save(path)
    cache = DetectionCache.load(path, device=dev)
    step = pipeline.RoIPathStep(shape, dev)
    d_host = step.to_device(batch)
    d_cache = dict(d_host)
    views = pipeline.cloud_inputs_from_cache(cache, names)
    st = cache.tags["RCNN"]
    for k, v in views.items():
        flat = {"gt_boxes": st.boxes, "gt_classes": st.classes, "scores": st.scores, "probs": st.probs}[k.split(".")[-1]]
        assert v.device.type == "cuda"
        lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * flat.element_size()
        assert v.numel() == 0 or lo <= v.data_ptr() < hi, f"{k} is not a view into the cache"
        assert torch.equal(v, d_host[k]), k
    d_cache.update(views)
    a = step.finalize(step.run_static(d_host, backward=False))
    b = step.finalize(step.run_static(d_cache, backward=False))
    for i in range(shape.images):
        for tag in ("RCNN", "RPN"):
            for x, y in zip(a["abc"][i][tag], b["abc"][i][tag]):
                if x is None:
                    assert y is None
                    continue
                for k in x:
                    assert torch.equal(x[k], y[k]), (i, tag, k)
    assert torch.equal(a["pooled_c"], b["pooled_c"])
    # the host-level mirror of the same hand-over: lookup -> process -> match_dual_teacher, nothing leaves the device
    inst = cache.lookup(names[0], "RCNN")
    on = integration.process(inst, inst.image_size, (shape.height, shape.width), "no")
    assert on.gt_boxes.tensor.device.type == "cuda" and on.gt_boxes.tensor.data_ptr() != inst.pred_boxes.tensor.data_ptr()
    want = coin_ref.process({"pred_boxes": per_file[names[0]]["RCNN"]["instances"].pred_boxes.tensor,
                             "pred_classes": per_file[names[0]]["RCNN"]["instances"].pred_classes,
                             "scores": per_file[names[0]]["RCNN"]["instances"].scores,
                             "probs": per_file[names[0]]["RCNN"]["instances"].probs}, inst.image_size,
                            (shape.height, shape.width))
    torch.testing.assert_close(on.gt_boxes.tensor.cpu(), want["gt_boxes"], rtol=1e-5, atol=1.2e-4)
    # a cache entry is never modified by the step (process clones; the reference deep-copies, gdino_collector.py:86)
    assert torch.equal(cache.lookup(names[0], "RCNN").pred_boxes.tensor.cpu(),
                       per_file[names[0]]["RCNN"]["instances"].pred_boxes.tensor)

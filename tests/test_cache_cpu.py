"""Collected-detections cache: reads a file laid out like the reference's GDINO_collect.pth
(pickled detectron2 Instances / Boxes inside the nested results dict) WITHOUT detectron2, keeps it as flat arrays,
round-trips through its own tensor-only file, and serves per-image lookups as views. Host logic only: runs on CPU."""
import os
import sys
import types

import pytest
import torch

from coin_b200.cache import DetectionCache, load_reference_results
from coin_b200.structures import Boxes, Instances


def _write_reference_file(path, per_file, key="results"):
    """torch.save of the reference's nested dict with objects whose classes pickle as detectron2.structures.*"""
    names = ("detectron2", "detectron2.structures", "detectron2.structures.instances", "detectron2.structures.boxes")
    mods = {n: types.ModuleType(n) for n in names}

    class FakeBoxes:                      # attribute layout of detectron2.structures.Boxes
        def __init__(self, tensor):
            self.tensor = tensor

    class FakeInstances:                  # attribute layout of detectron2.structures.Instances
        def __init__(self, image_size, **fields):
            self._image_size = image_size
            self._fields = dict(fields)

    FakeBoxes.__module__, FakeBoxes.__qualname__, FakeBoxes.__name__ = "detectron2.structures.boxes", "Boxes", "Boxes"
    FakeInstances.__module__, FakeInstances.__qualname__, FakeInstances.__name__ = "detectron2.structures.instances", "Instances", "Instances"
    mods["detectron2.structures.boxes"].Boxes = FakeBoxes
    mods["detectron2.structures.instances"].Instances = FakeInstances
    saved = {n: sys.modules.get(n) for n in names}
    sys.modules.update(mods)
    try:
        blob = {}
        for fname, rec in per_file.items():
            out = {k: v for k, v in rec.items() if k not in ("RCNN", "RPN")}
            for tag in ("RCNN", "RPN"):
                if tag in rec:
                    f = rec[tag]
                    out[tag] = {"instances": FakeInstances(f["image_size"], pred_boxes=FakeBoxes(f["boxes"]), scores=f["scores"],
                                                           pred_classes=f["classes"], probs=f["probs"])}
            blob[fname] = out
        torch.save({key: {"foggy_train": blob}, "iteration": -1}, path)
    finally:
        for n, m in saved.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m


def _dets(g, n, k1=9):
    xy = torch.rand(n, 2, generator=g) * 500
    wh = torch.rand(n, 2, generator=g) * 200 + 1
    probs = torch.softmax(torch.randn(n, k1, generator=g), dim=1)
    return {"image_size": (600, 1200), "boxes": torch.cat((xy, xy + wh), 1), "scores": probs.max(1).values,
            "classes": probs.argmax(1), "probs": probs}


def _same(inst: Instances, want: dict):
    assert inst.image_size == want["image_size"]
    assert isinstance(inst.pred_boxes, Boxes) and torch.equal(inst.pred_boxes.tensor, want["boxes"])
    assert torch.equal(inst.scores, want["scores"]) and torch.equal(inst.pred_classes, want["classes"])
    assert inst.pred_classes.dtype == torch.int64 and torch.equal(inst.probs, want["probs"])


def test_detection_cache_reads_reference_layout_and_round_trips(tmp_path):
    g = torch.Generator().manual_seed(3)
    per_file = {
        "a/img0.png": {"file_name": "a/img0.png", "image_id": 0, "height": 1024, "width": 2048, "RCNN": _dets(g, 17), "RPN": _dets(g, 5)},
        "a/img1.png": {"file_name": "a/img1.png", "image_id": "x1", "height": 1024, "width": 2048, "RCNN": _dets(g, 0)},
        "b/img2.png": {"file_name": "b/img2.png", "image_id": 2, "height": 720, "width": 1280, "RCNN": _dets(g, 100), "RPN": _dets(g, 31)},
    }
    ref_path = os.path.join(tmp_path, "GDINO_collect.pth")
    _write_reference_file(ref_path, per_file)
    assert "detectron2" not in sys.modules                       # the reader must not need it
    results = load_reference_results(ref_path)
    assert set(results) == {"foggy_train"} and isinstance(results["foggy_train"]["a/img0.png"]["RCNN"]["instances"], Instances)

    cache = DetectionCache.load_reference(ref_path)
    assert len(cache) == 3 and "b/img2.png" in cache and "nope.png" not in cache
    assert cache.sizes.tolist() == [[1024, 2048], [1024, 2048], [720, 1280]] and cache.image_ids == [0, "x1", 2]
    for name, rec in per_file.items():
        for tag in ("RCNN", "RPN"):
            assert cache.has_tag(name, tag) == (tag in rec)
            if tag in rec:
                _same(cache.lookup(name, tag), rec[tag])
    with pytest.raises(KeyError):
        cache.lookup("a/img1.png", "RPN")
    e = cache.entry("b/img2.png")
    assert e["height"] == 720 and e["image_id"] == 2 and set(e) == {"file_name", "image_id", "height", "width", "RCNN", "RPN"}
    # lookups are views into the flat arrays (no copy)
    v = cache.lookup("b/img2.png", "RCNN")
    assert v.scores.data_ptr() == cache.tags["RCNN"].scores[17:].data_ptr()

    # own format: tensors + lists only -> loads with weights_only=True
    own = os.path.join(tmp_path, "cache.pt")
    cache.save(own)
    again = DetectionCache.load(own)
    for name, rec in per_file.items():
        for tag in ("RCNN", "RPN"):
            if tag in rec:
                _same(again.lookup(name, tag), rec[tag])

    # update with another length (gdino_collector.py:93-101), visible at once, folded in by compact() / save()
    new = _dets(g, 7)
    cache.update("a/img1.png", "RPN", Instances(new["image_size"], pred_boxes=Boxes(new["boxes"]), scores=new["scores"],
                                                 pred_classes=new["classes"], probs=new["probs"]))
    _same(cache.lookup("a/img1.png", "RPN"), new)
    cache.compact()
    _same(cache.lookup("a/img1.png", "RPN"), new)
    _same(cache.lookup("a/img0.png", "RPN"), per_file["a/img0.png"]["RPN"])
    _same(cache.lookup("b/img2.png", "RCNN"), per_file["b/img2.png"]["RCNN"])
    cache.save(own)
    _same(DetectionCache.load(own).lookup("a/img1.png", "RPN"), new)
    with pytest.raises(KeyError):
        cache.update("unknown.png", "RCNN", cache.lookup("a/img0.png", "RCNN"))
    assert cache.to("cpu").nbytes() == cache.nbytes() > 0


def test_detection_cache_reads_trainer_checkpoint_key(tmp_path):
    """Trainer checkpoints carry the same dict under 'online_results' (trainer.py:252)."""
    g = torch.Generator().manual_seed(4)
    per_file = {"img.png": {"file_name": "img.png", "image_id": 9, "height": 600, "width": 800, "RCNN": _dets(g, 3)}}
    path = os.path.join(tmp_path, "model_0000999.pth")
    _write_reference_file(path, per_file, key="online_results")
    cache = DetectionCache.load_reference(path, "foggy_train")
    _same(cache.lookup("img.png"), per_file["img.png"]["RCNN"])
    with pytest.raises(ValueError):
        DetectionCache.from_state_dict({"format": "something else"})


def test_probs_presence_and_width_are_per_image_and_empty_instances_load():
    """images without a `probs` field (or a narrower one) must not come back with fabricated zero columns, and
    an Instances without any field (len() raises) is an image with no detections, not a load error."""
    def inst(n, k1):
        i = Instances((600, 1200))
        i.pred_boxes = Boxes(torch.rand(n, 4) * 100)
        i.scores = torch.rand(n)
        i.pred_classes = torch.randint(0, 8, (n,))
        if k1:
            i.probs = torch.rand(n, k1)
        return i
    per_file = {"a.png": {"file_name": "a.png", "image_id": 0, "height": 1024, "width": 2048, "RCNN": {"instances": inst(3, 9)}},
                "b.png": {"file_name": "b.png", "image_id": 1, "height": 1024, "width": 2048, "RCNN": {"instances": inst(2, 0)}},
                "c.png": {"file_name": "c.png", "image_id": 2, "height": 1024, "width": 2048, "RCNN": {"instances": inst(4, 5)}},
                "d.png": {"file_name": "d.png", "image_id": 3, "height": 1024, "width": 2048,
                          "RCNN": {"instances": Instances((600, 1200))}}}
    cache = DetectionCache.from_reference_dict(per_file)
    assert cache.lookup("a.png").probs.shape == (3, 9)
    assert not cache.lookup("b.png").has("probs")
    assert cache.lookup("c.png").probs.shape == (4, 5)
    assert torch.equal(cache.lookup("c.png").probs, per_file["c.png"]["RCNN"]["instances"].probs)
    assert len(cache.lookup("d.png").scores) == 0
    again = DetectionCache.from_state_dict(cache.state_dict())
    assert not again.lookup("b.png").has("probs") and again.lookup("c.png").probs.shape == (4, 5)


def test_reference_loader_refuses_unknown_globals(tmp_path):
    """a detections file is data; a pickle that names any other importable callable is refused."""
    import os
    import pickle
    path = os.path.join(tmp_path, "evil.pth")
    torch.save({"results": {"ds": {}}, "hook": os.path.join}, path)
    with pytest.raises(pickle.UnpicklingError):
        load_reference_results(path)

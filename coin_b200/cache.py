"""Device-resident structure-of-arrays cache of collected detections.

The reference keeps the cloud / CLIP detections of every training image in a nested dict of pickled
``Instances`` on the host (``coin/modeling/meta_arch/gdino_collector.py:51-101``, ``clip_collector.py:46-80``),
writes it with ``torch.save({'results': ...}, 'GDINO_collect.pth')`` (``coin/engine/pre_train.py:148-161``), reloads it
with ``torch.load`` (``coin/engine/trainer.py:231-232,252``), ``deepcopy``s an entry per lookup
(``gdino_collector.py:86``) and moves it to the GPU every step (``trainer.py:457-459``). Layout of that dict::

    results[dataset_name][file_name] = {'file_name', 'image_id', 'height', 'width',
                                        'RCNN': {'instances': Instances(pred_boxes, scores, pred_classes, probs)},
                                        'RPN':  {'instances': Instances(...)}}          # 'RPN' appears after update()

``DetectionCache`` holds the same information as a handful of flat tensors per tag (offsets + boxes + scores + classes +
probs) that live on the device, so the lookup feeding ``process`` (A13) and ``match_dual_teacher`` (A9) is a pair of
slices - no pickle, no deepcopy, no host-to-device copy. ``load_reference`` reads the reference's own ``.pth`` file
(the pickled detectron2 classes are mapped onto ``coin_b200.structures``; detectron2 is not needed), ``save`` / ``load``
use a plain tensor file that ``torch.load(weights_only=True)`` accepts.
"""
import io
import pickle
from typing import Any, Dict, Iterable, List, Optional, Tuple

import torch

from .structures import Boxes, Instances

TAGS = ("RCNN", "RPN", "RPN_AUG")
_FIELDS = ("pred_boxes", "scores", "pred_classes", "probs")
FORMAT_VERSION = 1


class _D2Unpickler(pickle.Unpickler):
    """Maps the two detectron2 classes inside the reference's files onto their mirrors (same attribute layout:
    ``Instances._image_size / _fields``, ``Boxes.tensor``). Everything else must be on a short allowlist - the tensor
    rebuild helpers ``torch.save`` emits and plain containers / scalars: a pickle can name ANY importable callable, and a
    collected-detections file is data, so an unknown global is refused instead of imported."""
    _MAP = {("detectron2.structures.instances", "Instances"): Instances,
            ("detectron2.structures.boxes", "Boxes"): Boxes,
            ("detectron2.structures", "Instances"): Instances,
            ("detectron2.structures", "Boxes"): Boxes}
    _ALLOWED = {
        "collections": {"OrderedDict", "defaultdict"},
        "builtins": {"dict", "list", "tuple", "set", "frozenset", "int", "float", "bool", "str", "bytes", "complex",
                     "slice", "range", "bytearray"},
        "torch._utils": {"_rebuild_tensor_v2", "_rebuild_tensor", "_rebuild_parameter", "_rebuild_qtensor"},
        "torch": {"Size", "device", "dtype", "float32", "float64", "float16", "bfloat16", "int64", "int32", "int16",
                  "int8", "uint8", "bool", "FloatStorage", "DoubleStorage", "HalfStorage", "BFloat16Storage",
                  "LongStorage", "IntStorage", "ShortStorage", "CharStorage", "ByteStorage", "BoolStorage"},
        "torch.storage": {"_load_from_bytes", "TypedStorage", "UntypedStorage"},
        "numpy": {"dtype", "ndarray"},
        "numpy.core.multiarray": {"_reconstruct", "scalar"},
        "numpy._core.multiarray": {"_reconstruct", "scalar"},
    }

    def find_class(self, module: str, name: str):
        hit = self._MAP.get((module, name))
        if hit is not None:
            return hit
        if name in self._ALLOWED.get(module, ()):
            return super().find_class(module, name)
        raise pickle.UnpicklingError(
            f"coin_b200.cache: refusing to load global {module}.{name} from a detections file "
            "(only tensors, plain containers and detectron2 Instances / Boxes are expected)")


class _D2Pickle:
    """The ``pickle_module`` protocol of ``torch.load``."""
    __name__ = "coin_b200.cache._D2Pickle"
    Unpickler = _D2Unpickler
    Pickler = pickle.Pickler
    HIGHEST_PROTOCOL = pickle.HIGHEST_PROTOCOL

    @staticmethod
    def load(f, **kw):
        return _D2Unpickler(f, **kw).load()

    @staticmethod
    def loads(b, **kw):
        return _D2Unpickler(io.BytesIO(b), **kw).load()

    dump = staticmethod(pickle.dump)
    dumps = staticmethod(pickle.dumps)


def load_reference_results(path: str) -> Dict[str, Dict[str, dict]]:
    """The ``results`` dict of a file written by the reference (``GDINO_collect.pth``, or a trainer checkpoint that
    carries ``online_results``), with ``coin_b200`` Instances / Boxes in place of detectron2's."""
    blob = torch.load(path, map_location="cpu", pickle_module=_D2Pickle, weights_only=False)
    for key in ("results", "online_results"):
        if isinstance(blob, dict) and key in blob and blob[key] is not None:
            return blob[key]
    raise KeyError(f"{path}: neither 'results' nor 'online_results' found")


class _TagStore:
    """Flat arrays of one tag: image i owns rows offsets[i] : offsets[i + 1]."""

    def __init__(self, offsets, boxes, scores, classes, probs, image_sizes, present, probs_width=None):
        self.offsets, self.boxes, self.scores, self.classes, self.probs = offsets, boxes, scores, classes, probs
        self.image_sizes, self.present = image_sizes, present   # [n, 2] (h, w) of the Instances; [n] bool
        # [n] width of the image's own `probs` field (0: the Instances had none): the flat array is padded to the widest
        # one, and a lookup must not hand out columns (or a whole field) the collector never wrote
        self.probs_width = (probs_width if probs_width is not None
                            else torch.where(present, int(probs.shape[1]), 0).to(torch.int64))

    def to(self, device) -> "_TagStore":
        return _TagStore(self.offsets, self.boxes.to(device), self.scores.to(device), self.classes.to(device),
                         self.probs.to(device), self.image_sizes, self.present, self.probs_width)

    def state(self) -> Dict[str, torch.Tensor]:
        return {"offsets": self.offsets, "boxes": self.boxes.cpu(), "scores": self.scores.cpu(), "classes": self.classes.cpu(),
                "probs": self.probs.cpu(), "image_sizes": self.image_sizes, "present": self.present,
                "probs_width": self.probs_width}

    @classmethod
    def from_state(cls, st: Dict[str, torch.Tensor]) -> "_TagStore":
        return cls(st["offsets"], st["boxes"], st["scores"], st["classes"], st["probs"], st["image_sizes"], st["present"],
                   st.get("probs_width"))


class DetectionCache:
    """SoA store of the collected detections of ONE dataset split; see the module docstring."""

    def __init__(self, file_names: List[str], image_ids: List[Any], sizes: torch.Tensor, tags: Dict[str, _TagStore]):
        self.file_names = list(file_names)
        self.image_ids = list(image_ids)
        self.sizes = sizes                      # int64 [n, 2]: original (height, width) of the image
        self.tags = tags
        self._index = {name: i for i, name in enumerate(self.file_names)}
        self._overlay: Dict[Tuple[str, str], Instances] = {}   # update()s since the last compact()

    # ---- construction ---------------------------------------------------------------------------------------------
    @classmethod
    def from_reference_dict(cls, per_file: Dict[str, dict], device=None) -> "DetectionCache":
        """``per_file`` = ``results[dataset_name]`` of the reference (values as written by the collectors)."""
        names = list(per_file.keys())
        n = len(names)
        sizes = torch.tensor([[int(per_file[k].get("height", 0)), int(per_file[k].get("width", 0))] for k in names],
                             dtype=torch.int64).reshape(n, 2)
        ids = [per_file[k].get("image_id") for k in names]
        tags: Dict[str, _TagStore] = {}
        for tag in TAGS:
            insts = [per_file[k][tag]["instances"] if tag in per_file[k] and "instances" in per_file[k][tag] else None
                     for k in names]
            if not any(i is not None for i in insts):
                continue
            tags[tag] = cls._pack(insts)
        cache = cls(names, ids, sizes, tags)
        return cache.to(device) if device is not None else cache

    @staticmethod
    def _pack(insts: List[Optional[Instances]]) -> _TagStore:
        k1 = 0
        for i in insts:
            if i is not None and i.has("probs") and i.probs.dim() == 2:
                k1 = max(k1, i.probs.shape[1])
        def _len(i):      # an Instances without any field has no length (len() raises): it holds no detections
            return 0 if i is None or not i.get_fields() else len(i)
        counts = [_len(i) for i in insts]
        offsets = torch.zeros(len(insts) + 1, dtype=torch.int64)
        offsets[1:] = torch.tensor(counts, dtype=torch.int64).cumsum(0) if insts else offsets[1:]
        total = int(offsets[-1])
        boxes = torch.zeros((total, 4), dtype=torch.float32)
        scores = torch.zeros((total,), dtype=torch.float32)
        classes = torch.zeros((total,), dtype=torch.int64)
        probs = torch.zeros((total, k1), dtype=torch.float32)
        image_sizes = torch.zeros((len(insts), 2), dtype=torch.int64)
        present = torch.zeros((len(insts),), dtype=torch.bool)
        probs_width = torch.zeros((len(insts),), dtype=torch.int64)
        for j, inst in enumerate(insts):
            if inst is None:
                continue
            present[j] = True
            image_sizes[j] = torch.tensor([int(inst.image_size[0]), int(inst.image_size[1])])
            if not inst.get_fields():
                continue
            a, b = int(offsets[j]), int(offsets[j + 1])
            if inst.has("probs") and inst.probs.dim() == 2:
                probs_width[j] = int(inst.probs.shape[1])
            if b == a:
                continue
            bx = inst.pred_boxes.tensor if isinstance(inst.pred_boxes, Boxes) else inst.pred_boxes
            boxes[a:b] = bx.detach().to("cpu", torch.float32)
            scores[a:b] = inst.scores.detach().to("cpu", torch.float32)
            classes[a:b] = inst.pred_classes.detach().to("cpu", torch.int64)
            if int(probs_width[j]):
                probs[a:b, : inst.probs.shape[1]] = inst.probs.detach().to("cpu", torch.float32)
        return _TagStore(offsets, boxes, scores, classes, probs, image_sizes, present, probs_width)

    @classmethod
    def load_reference(cls, path: str, dataset_name: Optional[str] = None, device=None) -> "DetectionCache":
        """Reads a file written by the reference itself (``GDINO_collect.pth`` or a checkpoint with ``online_results``)."""
        results = load_reference_results(path)
        if dataset_name is None:
            if len(results) != 1:
                raise ValueError(f"{path} holds {sorted(results)}: pass dataset_name")
            dataset_name = next(iter(results))
        return cls.from_reference_dict(results[dataset_name], device)

    # ---- persistence (plain tensors: loadable with weights_only=True) ------------------------------------------------
    def state_dict(self) -> Dict[str, Any]:
        self.compact()
        return {"format": "coin_b200.DetectionCache", "version": FORMAT_VERSION, "file_names": self.file_names,
                "image_ids": self.image_ids, "sizes": self.sizes, "tags": {t: s.state() for t, s in self.tags.items()}}

    def save(self, path: str) -> None:
        torch.save(self.state_dict(), path)

    @classmethod
    def from_state_dict(cls, st: Dict[str, Any], device=None) -> "DetectionCache":
        if st.get("format") != "coin_b200.DetectionCache" or st.get("version") != FORMAT_VERSION:
            raise ValueError("not a coin_b200 DetectionCache file (or an unknown version)")
        cache = cls(st["file_names"], st["image_ids"], st["sizes"], {t: _TagStore.from_state(s) for t, s in st["tags"].items()})
        return cache.to(device) if device is not None else cache

    @classmethod
    def load(cls, path: str, device=None) -> "DetectionCache":
        return cls.from_state_dict(torch.load(path, map_location="cpu", weights_only=True), device)

    # ---- use ---------------------------------------------------------------------------------------------------------
    def to(self, device) -> "DetectionCache":
        out = DetectionCache(self.file_names, self.image_ids, self.sizes, {t: s.to(device) for t, s in self.tags.items()})
        out._overlay = {k: v.to(device) for k, v in self._overlay.items()}
        return out

    def __len__(self) -> int:
        return len(self.file_names)

    def __contains__(self, file_name: str) -> bool:
        return file_name in self._index

    def has_tag(self, file_name: str, tag: str) -> bool:
        if (file_name, tag) in self._overlay:
            return True
        return tag in self.tags and file_name in self._index and bool(self.tags[tag].present[self._index[file_name]])

    def lookup(self, file_name: str, tag: str = "RCNN") -> Instances:
        """The detections of one image as ``Instances`` whose fields are VIEWS into the flat device arrays (what
        ``GDINO_COLLECTOR.forward`` returns after its deepcopy; callers that modify boxes in place must clone, as
        ``BASE_Trainer.process`` (base.py:80-126) does through ``Boxes.scale``)."""
        hit = self._overlay.get((file_name, tag))
        if hit is not None:
            return hit
        i = self._index[file_name]
        st = self.tags[tag]
        if not bool(st.present[i]):
            raise KeyError(f"{file_name} has no '{tag}' detections")
        a, b = int(st.offsets[i]), int(st.offsets[i + 1])
        inst = Instances((int(st.image_sizes[i, 0]), int(st.image_sizes[i, 1])))
        inst.pred_boxes = Boxes(st.boxes[a:b])
        inst.scores = st.scores[a:b]
        inst.pred_classes = st.classes[a:b]
        kw = int(st.probs_width[i])
        if kw:
            inst.probs = st.probs[a:b] if kw == st.probs.shape[1] else st.probs[a:b, :kw]
        return inst

    def entry(self, file_name: str) -> Dict[str, Any]:
        """The reference's per-file dict (``GDINO_COLLECTOR.forward(file_name)``), fields as views."""
        i = self._index[file_name]
        out = {"file_name": file_name, "image_id": self.image_ids[i], "height": int(self.sizes[i, 0]), "width": int(self.sizes[i, 1])}
        for tag in TAGS:
            if self.has_tag(file_name, tag):
                out[tag] = {"instances": self.lookup(file_name, tag)}
        return out

    def update(self, file_name: str, tag: str, instances: Instances) -> None:
        """``GDINO_COLLECTOR.update`` (gdino_collector.py:93-101): replaces one image's detections. The new set may have
        a different length, so it is kept beside the flat arrays until ``compact()`` folds it in."""
        if file_name not in self._index:
            raise KeyError(file_name)
        self._overlay[(file_name, tag)] = instances

    def compact(self) -> None:
        """Folds the pending updates into the flat arrays (one rebuild per tag that was touched)."""
        if not self._overlay:
            return
        touched = {tag for (_, tag) in self._overlay}
        for tag in touched:
            device = self.tags[tag].boxes.device if tag in self.tags else next(iter(self._overlay.values())).scores.device
            insts: List[Optional[Instances]] = []
            for name in self.file_names:
                if (name, tag) in self._overlay:
                    insts.append(self._overlay[(name, tag)].to("cpu"))
                elif tag in self.tags and bool(self.tags[tag].present[self._index[name]]):
                    insts.append(self._lookup_cpu(name, tag))
                else:
                    insts.append(None)
            self.tags[tag] = self._pack(insts).to(device)
        self._overlay.clear()

    def _lookup_cpu(self, file_name: str, tag: str) -> Instances:
        i = self._index[file_name]
        st = self.tags[tag]
        a, b = int(st.offsets[i]), int(st.offsets[i + 1])
        inst = Instances((int(st.image_sizes[i, 0]), int(st.image_sizes[i, 1])))
        inst.pred_boxes = Boxes(st.boxes[a:b].cpu())
        inst.scores = st.scores[a:b].cpu()
        inst.pred_classes = st.classes[a:b].cpu()
        kw = int(st.probs_width[i])
        if kw:
            inst.probs = st.probs[a:b, :kw].cpu()
        return inst

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for s in self.tags.values() for t in (s.boxes, s.scores, s.classes, s.probs))

    def file_names_iter(self) -> Iterable[str]:
        return iter(self.file_names)

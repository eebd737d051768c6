"""Runs the reference's OWN Python, unmodified, in the build container (TEST INFRASTRUCTURE ONLY).

Every hot-path module of /root/reference imports detectron2 / fvcore / supervision / groundingdino at
module top, and none of those is installable here. ``install()`` puts a meta-path finder in front of
the import system that resolves

  * ``coin.<module>`` for the modules listed in REAL_COIN: the reference's file, loaded BY PATH from
    /root/reference, byte for byte (nothing is copied into this repository);
  * ``detectron2.<module>`` for which oracle/d2_shim holds a file: a minimal restatement of that
    detectron2-0.5 module (Boxes, Instances, pairwise_iou, Matcher, Box2BoxTransform, batched_nms,
    subsample_labels, add_ground_truth_to_proposals, retry_if_cuda_oom ...);
  * every other ``detectron2.* / fvcore.* / supervision / groundingdino.* / coin.* ...`` name: an
    inert stub module whose attributes are inert stub classes (usable as base class, decorator,
    registry, logger).

So ``CoinTrainer.match_dual_teacher``, ``delete_duplicate_boxes``, ``online_boxes_merging``,
``BASE_Trainer.process``, ``fast_rcnn_inference_single_image``, ``label_and_sample_proposals``,
``label_and_sample_anchors``, ``GDINO_PROCESSOR.nms`` and ``GDINO.resize_boxes`` execute exactly as
the reference wrote them; tests/golden/make_golden_ref.py freezes their outputs. Nothing here is
importable from coin_b200, and /root/reference does not exist on the GPU box: only the frozen
fixtures travel.
"""
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("COIN_REFERENCE_ROOT", "/root/reference")
SHIM_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "d2_shim")

# reference modules executed for real (everything else under coin.* is a stub)
REAL_COIN = {
    "coin.utils.util": "coin/utils/util.py",
    "coin.utils.losses": "coin/utils/losses.py",
    "coin.layers.nms": "coin/layers/nms.py",
    "coin.engine.base": "coin/engine/base.py",
    "coin.engine.trainer": "coin/engine/trainer.py",
    "coin.modeling.utils": "coin/modeling/utils.py",
    "coin.modeling.roi_heads.fast_rcnn": "coin/modeling/roi_heads/fast_rcnn.py",
    "coin.modeling.roi_heads.clip_roi_heads": "coin/modeling/roi_heads/clip_roi_heads.py",
    "coin.modeling.proposal_generator.rpn": "coin/modeling/proposal_generator/rpn.py",
    "coin.modeling.meta_arch.gdino_processor": "coin/modeling/meta_arch/gdino_processor.py",
    "coin.modeling.meta_arch.gdino": "coin/modeling/meta_arch/gdino.py",
    "coin.evaluation.cloud_pascal_voc_evaluation": "coin/evaluation/cloud_pascal_voc_evaluation.py",
}
STUB_TOPS = {"detectron2", "fvcore", "supervision", "groundingdino", "maskrcnn_benchmark", "iopath", "yacs",
             "pycocotools", "tensorboardX", "omegaconf", "termcolor", "lvis", "coin", "clip", "ftfy", "timm"}


class StubMeta(type):
    """Class of the inert stand-ins: any attribute is another stub; calling a stub with one callable
    returns that callable (decorator / registry use), anything else returns a stub. Classes that merely
    INHERIT from a stub (the reference's own classes) instantiate normally."""

    def __getattr__(cls, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return make_stub(f"{cls.__name__}.{name}")

    def __call__(cls, *args, **kwargs):
        if cls.__dict__.get("_coin_stub", False):
            if len(args) == 1 and not kwargs and callable(args[0]) and not isinstance(args[0], StubMeta):
                return args[0]
            return make_stub(cls.__name__ + "()")
        return super().__call__(*args, **kwargs)

    def __iter__(cls):
        return iter(())

    def __bool__(cls):
        return False

    def __ge__(cls, other):
        return True

    def __le__(cls, other):
        return True

    __gt__ = __ge__
    __lt__ = __le__


def make_stub(name: str):
    return StubMeta(name.replace(".", "_"), (), {"_coin_stub": True, "__module__": "oracle.ref_loader"})


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        stub = make_stub(f"{self.__name__}.{name}")
        setattr(self, name, stub)
        return stub


class _StubLoader(importlib.abc.Loader):
    def create_module(self, spec):
        mod = _StubModule(spec.name)
        mod.__path__ = []          # a package: sub-modules resolve through the finder again
        return mod

    def exec_module(self, module):
        pass


class _ShimLoader(importlib.machinery.SourceFileLoader):
    """A real file of oracle/d2_shim; names it does not define fall through to stubs."""

    def exec_module(self, module):
        super().exec_module(module)
        name = module.__name__

        def fallthrough(attr, _name=name):
            if attr.startswith("__") and attr.endswith("__"):
                raise AttributeError(attr)
            return make_stub(f"{_name}.{attr}")
        module.__getattr__ = fallthrough


class RefFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        top = fullname.split(".", 1)[0]
        if fullname in REAL_COIN:
            file = os.path.join(REF_ROOT, REAL_COIN[fullname])
            return importlib.util.spec_from_file_location(fullname, file)
        if top not in STUB_TOPS:
            return None
        rel = fullname.replace(".", os.sep)
        for cand, is_pkg in ((os.path.join(SHIM_ROOT, rel, "__init__.py"), True), (os.path.join(SHIM_ROOT, rel + ".py"), False)):
            if os.path.exists(cand):
                return importlib.util.spec_from_file_location(
                    fullname, cand, loader=_ShimLoader(fullname, cand),
                    submodule_search_locations=[] if is_pkg else None)
        return importlib.machinery.ModuleSpec(fullname, _StubLoader(), is_package=True)


_FINDER = None


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "coin"))


def install() -> None:
    """Idempotent. Raises when the reference tree is absent (e.g. on the GPU box)."""
    global _FINDER
    if not available():
        raise RuntimeError(f"{REF_ROOT} is not present: the reference can only be executed in the build container")
    if _FINDER is None:
        repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        if repo not in sys.path:
            sys.path.insert(0, repo)
        _FINDER = RefFinder()
        sys.meta_path.insert(0, _FINDER)


def load(name: str):
    """import_module of a reference module (``coin.utils.util`` ...) under the stub finder."""
    install()
    return importlib.import_module(name)

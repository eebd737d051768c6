"""CPU oracle for the COIN RoI / box-decode / IoU-match / NMS path.

TEST INFRASTRUCTURE ONLY. Nothing under ``coin_b200/`` may import this package: only ``tests/``,
``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py`` use it, as the checker (or, in
``bench.py --impl reference`` / ``cpu_baseline``, as the thing timed on the host cores).

Layout
------
scalar_ref.c   plain-C restatement of torchvision's roi_align fwd/bwd and nms and of detectron2's
               pairwise_iou (built by ``make -C oracle`` into ``oracle/_build/liboracle.so``)
clib.py        ctypes loader for the above
d2_ref.py      restatement (torch CPU ops) of the detectron2-0.5 operators the reference imports
coin_ref.py    restatement of the arithmetic the reference authors itself
               (coin/layers/nms.py, coin/engine/trainer.py:338-485, coin/utils/util.py:434-507,
               coin/engine/base.py:80-136, coin/modeling/roi_heads/fast_rcnn.py:116-175, ...)
d2_shim/       a stub ``detectron2.layers.batched_nms`` so the reference's own coin/layers/nms.py can
               be loaded, unmodified, by tests/golden/make_golden.py in the build container

Parity status: the reference repository has no tests and no golden vectors, so
the oracle is pinned against (a) the torchvision 0.26 CPU operators of this image - the same
algorithms the reference reaches through detectron2 0.5 / torchvision 0.10.1 - and (b) outputs of
the reference's own ``coin/layers/nms.py`` executed in the build container and frozen under
``tests/golden/`` together with the generating script. The detectron2-0.5 pieces (Matcher,
Box2BoxTransform, ROIPooler level rule) and ``match_dual_teacher`` cannot be executed here
(detectron2 is not installable offline): for those the oracle is "parity unpinned" against the
reference binary and is checked by properties and hand-computed cases only.
"""

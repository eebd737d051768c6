"""Stub package: ONLY provides detectron2.layers.batched_nms so that the reference's
coin/layers/nms.py can be imported unmodified by tests/golden/make_golden.py (build container only).
TEST INFRASTRUCTURE; never imported by coin_b200."""

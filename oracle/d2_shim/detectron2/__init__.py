"""Minimal stand-in for detectron2 0.5 (docs/Environment.md:51 of the reference pins that release; it is
not installable offline). Only what the COIN hot path touches is restated here, from the published
0.5 sources, so that the reference's own modules can be executed unmodified by oracle/ref_loader.py.
TEST INFRASTRUCTURE; never imported by coin_b200. Names not defined here resolve to inert stubs."""
__version__ = "0.5-shim"

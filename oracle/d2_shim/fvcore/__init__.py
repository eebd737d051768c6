"""Minimal stand-in for fvcore (0.1.5.post20221221 per docs/Environment.md:102 of the reference)."""

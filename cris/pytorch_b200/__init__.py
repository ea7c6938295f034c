"""cris.pytorch_b200 — B200-native (sm_100a) drop-in for the CRIS forward/backward hot path.

Public surface = the reference's: `CRIS(cfg)` (model/segmenter.py:10) and `build_segmenter(args)`
(model/__init__.py:32).  Kernels live in csrc/ behind the C ABI of include/cris_b200.h.
"""
from .module import CRIS, build_segmenter  # noqa: F401

__all__ = ["CRIS", "build_segmenter"]

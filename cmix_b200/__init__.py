"""cmix_b200 — B200-native per-bit context-mixing predictor behind cmix's Predictor surface.

The product is the C-ABI shared library built from cmix_b200/csrc (see include/cmixb200.h);
this package is the thin Python binding used by tests/ and bench.py. It never falls back to
a CPU implementation: importing it without the built library raises.
"""
from .capi import Predictor, load_library, build_library, N_EXT, LIB_PATH  # noqa: F401

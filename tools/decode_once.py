#!/usr/bin/env python
"""Device encoder -> device decoder over n bytes of the golden text (for ncu launch lists of the decode loop):  python tools/decode_once.py [n_bytes]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CMIXB200_PPMD_MB", "512")
import numpy as np
import cmix_b200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
g = np.load(os.path.join(ROOT, "tests", "golden", "full_text.npz"))
stream = g["stream"][:n]
enc = cmix_b200.Predictor(g["vocab"]); enc.coder_begin(2 * n + 64); enc.code_bytes(stream); arch = enc.coder_finish(); enc.close()
dec = cmix_b200.Predictor(g["vocab"])
out = dec.decode_bytes(arch, n)
print("round trip", out.tobytes() == stream.tobytes())
dec.close()

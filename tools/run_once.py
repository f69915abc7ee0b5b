#!/usr/bin/env python
"""One resident predictor over n bytes of the bench text (for ncu / compute-sanitizer captures):  python tools/run_once.py [n_bytes]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
os.environ.setdefault("CMIXB200_PPMD_MB", "1024")
import numpy as np
import cmix_b200
from gen_synth import synth_text

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
text = np.frombuffer(synth_text(max(n, 16384), 0xE9E80001), dtype=np.uint8).copy()
vocab = np.zeros(256, dtype=np.uint8); vocab[np.unique(text)] = 1
P = cmix_b200.Predictor(vocab)
p = P.code_bytes(text[:n])
print("coded", n, "bytes; mean p", float(p.mean()))
P.close()

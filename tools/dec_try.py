import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
os.environ.setdefault("CMIXB200_PPMD_MB", "512")
import cmix_b200
g = np.load('/root/repo/tests/golden/full_text.npz')
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
stream = g["stream"][:n]
enc = cmix_b200.Predictor(g["vocab"]); enc.coder_begin(2*n+64); enc.code_bytes(stream); arch = enc.coder_finish(); enc.close()
dec = cmix_b200.Predictor(g["vocab"])
out = dec.decode_bytes(arch, n)
print("ok", out.tobytes() == stream.tobytes())

import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
os.environ.setdefault("CMIXB200_PPMD_MB", "512")
import cmix_b200
g = np.load('/root/repo/tests/golden/full_text.npz')
bits = np.unpackbits(g["stream"])
P = cmix_b200.Predictor(g["vocab"])
fc = g["first_codes"]
for t in range(64):
    p = P.Predict()
    codes = P.debug_fetch(11, (2022,), np.uint16)
    okp = p == g["p"][t]
    bad = np.nonzero(codes != fc[t])[0]
    if not okp or bad.size:
        print("bit", t, "p ok", okp, "bad code slots", bad[:12], "n", bad.size)
        break
    P.Perceive(int(bits[t]))
else:
    print("64 bits fine")

# per-phase cycle profile of the mix kernel (clock64 counters) on synthetic replay streams
import sys, time, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import cmix_b200
from conftest import synthetic_streams
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
stream, vocab, codes, ppmd = synthetic_streams(n, seed=1)
P = cmix_b200.Predictor(vocab)
P.code_bytes(stream[:64], codes[:512], ppmd[:64])       # warm-up
P.debug_fetch(6, (32,), np.uint64)                        # enable counters
t0 = time.time()
P.code_bytes(stream[64:], codes[512:], ppmd[64:])
dt = time.time() - t0
prof = P.debug_fetch(6, (32,), np.uint64).astype(np.float64)
bits = (n - 64) * 8
names = ['stage', 'aux+slots', 'row switch', 'chains', 'sync1', 'final+coeff', 'sync2', 'update']
print('wall %.3f s for %d bits = %.2f us/bit' % (dt, bits, dt / bits * 1e6))
for i, nm in enumerate(names):
    print('%-12s %8.0f cycles/bit' % (nm, prof[i] / bits))
print('total        %8.0f cycles/bit' % (prof[:8].sum() / bits))

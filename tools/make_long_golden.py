"""Long-run whole-predictor fixture from a level-0 dump of the UNMODIFIED reference (build container only).

    python tools/make_long_golden.py <dump prefix> <name>

tests/golden/<name>.npz: the coded stream (>= 256 KiB of tools/gen_synth.py text), the vocabulary, one CRC32 per 4096 coded
bits over the bit patterns of Predictor::Predict(), the first and last 4096 probabilities and the reference's cross entropy.
Recipe (tests/test_long_run.py):
    python -c "from gen_synth import synth_text; open('t256k.txt','wb').write(synth_text(262144))"
    oracle/_ref/oracle_dump dump n t256k.txt <prefix> 0
"""
import os, sys, zlib
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from oracle_io import Dump  # noqa: E402


def main():
    prefix, name = sys.argv[1], sys.argv[2]
    d = Dump(prefix)
    n = d.n_bytes - d.n_bytes % 512
    nb = n * 8
    p = np.ascontiguousarray(d.p[:nb], dtype=np.float32)
    crc = np.array([zlib.crc32(p[b:b + 4096].tobytes()) for b in range(0, nb, 4096)], dtype=np.uint32)
    bits = np.unpackbits(np.asarray(d.stream[:n], dtype=np.uint8))
    pd = p.astype(np.float64)
    bpc = float(-np.log2(np.where(bits == 1, pd, 1 - pd).clip(1.0 / 65536, 1)).sum() / n)
    out = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(out, stream=d.stream[:n], vocab=d.vocab, crc_p=crc, p_head=p[:4096], p_tail=p[-4096:], bpc=np.array([bpc]))
    print(name, n, "bytes, bpc %.4f ->" % bpc, os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    main()

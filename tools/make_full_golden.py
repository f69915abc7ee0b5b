"""Whole-predictor fixtures from dumps of the UNMODIFIED reference (build container only).

    python tools/make_full_golden.py <dump prefix> <name> [n_bytes]

tests/golden/<name>.npz: the coded stream, Predictor::Predict() of every bit, and one CRC32 per 4096 coded bits over the
reference's 431 FXCM codes and over its 1591 PAQ8 codes. With every model group resident these pin the complete path
(no replayed inputs): tests/test_full_predictor.py. Recipes (tools/gen_synth.py, default seeds):
    full_text   gen_synth text 40000 | head -c 12000;  oracle_dump dump n ...
    full_bin    gen_synth binary 140000;               oracle_dump dump c ... 1 20000
    full_wrt    gen_synth text 40000;                  oracle_dump dump c ... 1 40000 english.dic
"""
import os, sys, zlib
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from oracle_io import Dump  # noqa: E402


def main():
    prefix, name = sys.argv[1], sys.argv[2]
    d = Dump(prefix)
    n = min(int(sys.argv[3]) if len(sys.argv) > 3 else d.n_bytes, d.n_bytes)
    n -= n % 512
    nb = n * 8
    ext = np.memmap(prefix + ".ext.u16", dtype=np.uint16, mode="r").reshape(-1, 2022)
    crc_fx = np.array([zlib.crc32(np.ascontiguousarray(ext[b:b + 4096, :431]).tobytes()) for b in range(0, nb, 4096)], dtype=np.uint32)
    crc_p8 = np.array([zlib.crc32(np.ascontiguousarray(ext[b:b + 4096, 431:]).tobytes()) for b in range(0, nb, 4096)], dtype=np.uint32)
    out = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(out, stream=d.stream[:n], vocab=d.vocab, p=d.p[:nb], crc_fx=crc_fx, crc_p8=crc_p8, first_codes=np.ascontiguousarray(ext[:64]),
                        mode=np.array([d.meta["mode"]]), dictionary=np.array([int(d.meta["dictionary"])]))
    print(name, n, "bytes ->", os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    main()

"""Fixtures for the resident FXCM model from dumps of the UNMODIFIED reference (build container only).

    python tools/make_fxcm_golden.py <dump prefix> <name> [n_bytes]

<dump prefix> is an oracle/_ref/oracle_dump level>=1 dump (see oracle/ref_driver.cpp). The fixture
tests/golden/<name>.npz holds the coded stream, the LSTM feedback FXCM consumed per bit (lstmpr | lstmex << 16)
and one CRC32 per 4096 coded bits over the reference's 431 FXCM codes of every bit - enough to pin the model
bit for bit on the CPU (tools/fxcm_check.cpp) and on the GPU without shipping 862 bytes per bit.
Recipes of the committed fixtures (tools/gen_synth.py generators, seeds as in the command lines):
    fxcm_text    gen_synth text 40000 | head -c 12000; oracle_dump dump n ...                   (no preprocessing)
    fxcm_bin     gen_synth binary 140000;              oracle_dump dump c ... 1 20000            (EXE/JPEG/DEFAULT blocks)
    fxcm_wrt     gen_synth text 40000;                 oracle_dump dump c ... 1 40000 english.dic (WRT + Pretrain)
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
    n -= n % 512                                   # whole CRC blocks
    nb = n * 8
    fx = np.fromfile(prefix + ".lstmfx.u32", dtype=np.uint32)[:nb]
    ext = np.memmap(prefix + ".ext.u16", dtype=np.uint16, mode="r").reshape(-1, 2022)
    crc = np.array([zlib.crc32(np.ascontiguousarray(ext[b:b + 4096, :431]).tobytes()) for b in range(0, nb, 4096)], dtype=np.uint32)
    out = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(out, stream=d.stream[:n], vocab=d.vocab, lstmfx=fx, crc=crc, first_codes=np.ascontiguousarray(ext[:64, :431]),
                        mode=np.array([d.meta["mode"]]), dictionary=np.array([int(d.meta["dictionary"])]))
    print(name, n, "bytes ->", os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    main()

"""Make tests/golden/ppmd_*.npz from dumps of the unmodified reference.

    oracle/_ref/oracle_dump dump n <file> <prefix> 1 <n_bytes>     # writes <prefix>.ppmd.f32 (+ .stream, .meta)
    python tools/make_ppmd_golden.py <prefix> tests/golden/ppmd_<name>.npz

The fixture holds the coded stream, the vocabulary and one CRC-32 per byte of the 256-float distribution
PPMD::ByteUpdate leaves after that byte (reference src/models/ppmd.cpp:1328-1338) - 4 bytes instead of 1 KB.
"""
import sys
import zlib

import numpy as np


def main():
    prefix, out = sys.argv[1], sys.argv[2]
    meta = dict(line.split(None, 1) for line in open(prefix + ".meta"))
    n = int(meta["n_bytes"])
    vocab = np.array([int(c) for c in meta["vocab"].strip()], dtype=np.uint8)
    stream = np.fromfile(prefix + ".stream", dtype=np.uint8)[:n]
    pp = np.fromfile(prefix + ".ppmd.f32", dtype=np.float32).reshape(n, 256)
    crc = np.array([zlib.crc32(pp[t].tobytes()) for t in range(n)], dtype=np.uint32)
    np.savez_compressed(out, stream=stream, vocab=vocab, crc=crc)
    print(out, n, "bytes")


if __name__ == "__main__":
    main()

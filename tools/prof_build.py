#!/usr/bin/env python
"""Build a profiling variant of the library (-DP8_PROF -DFX_PROF: per-phase cycle counters in the producer kernels) next to
the product library and print where a bit's time goes. Never used by the product path, the tests or bench.py.

    python tools/prof_build.py build            # here (nvcc cross-compiles)
    python tools/prof_build.py run [n_bytes]    # on the GPU box: CMIXB200_LIB=<prof lib> is set by this script
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "cmix_b200", "csrc")
PROF_LIB = os.path.join(CSRC, "libcmixb200_prof.so")


def build():
    sys.path.insert(0, ROOT)
    from cmix_b200.capi import NVCC_COMPILE, NVCC_LINK
    objs, jobs = [], []
    for u in ["engine.cu", "fxcm_dev.cu", "paq8_dev.cu"]:
        obj = os.path.join(CSRC, u[:-3] + "_prof.o")
        objs.append(obj)
        jobs.append(subprocess.Popen(["nvcc"] + NVCC_COMPILE + ["-DP8_PROF", "-DFX_PROF", "-c", os.path.join(CSRC, u), "-o", obj]))
    if any(j.wait() != 0 for j in jobs):
        raise SystemExit("nvcc failed")
    subprocess.run(["nvcc"] + NVCC_LINK + objs + ["-o", PROF_LIB], check=True)
    print(PROF_LIB)


def run(n_bytes):
    os.environ["CMIXB200_LIB"] = PROF_LIB
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import numpy as np
    import torch
    import cmix_b200
    from cmix_b200.capi import load_library
    from gen_synth import synth_text
    text = np.frombuffer(synth_text(n_bytes * 2, 0xE9E80001), dtype=np.uint8).copy()
    vocab = np.zeros(256, dtype=np.uint8)
    vocab[np.unique(text)] = 1
    P = cmix_b200.Predictor(vocab)
    P.code_bytes(text[:n_bytes])
    lib = load_library()
    sm_mhz = torch.cuda.get_device_properties(0).clock_rate / 1e3 if hasattr(torch.cuda.get_device_properties(0), "clock_rate") else 1965.0
    for name, fn, rows in (("paq8", "cmixb200_p8_prof", 96), ("fxcm", "cmixb200_fx_prof", 24)):
        if not hasattr(lib, fn):
            continue
        buf = (ctypes.c_ulonglong * (2 * rows))()
        getattr(lib, fn)(buf, 1)
    P.time_mix_kernel(True)
    P.code_bytes(text[n_bytes:2 * n_bytes])
    for w, k in enumerate(["mix", "small", "lstm", "ppmd", "fxcm", "paq8"]):
        ms, n = P.kernel_ms(w)
        print("%-6s %8.2f us/bit (%d launches)" % (k, ms * 1e3 / (n_bytes * 8), n))
    for name, fn, rows in (("paq8", "cmixb200_p8_prof", 96), ("fxcm", "cmixb200_fx_prof", 24)):
        try:
            f = getattr(lib, fn)
        except AttributeError:
            continue
        buf = (ctypes.c_ulonglong * (2 * rows))()
        f(buf, 0)
        a = np.array(list(buf), dtype=np.float64).reshape(2, rows)
        print("%s: cycles per bit by phase (byte-boundary bits | other bits); clock %.0f MHz" % (name, sm_mhz))
        for k in range(rows):
            if a[0, k] or a[1, k]:
                print("  phase %2d  %9.0f | %9.0f" % (k, a[0, k] / n_bytes, a[1, k] / (7 * n_bytes)))
        print("  total     %9.0f | %9.0f   -> %.1f us/bit average" % (a[0, :24].sum() / n_bytes, a[1, :24].sum() / (7 * n_bytes), a[:, :24].sum() / (8 * n_bytes) / sm_mhz))
    P.close()


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    else:
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 1024)

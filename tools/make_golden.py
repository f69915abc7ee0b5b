"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    python tools/make_golden.py

Needs oracle/_ref/oracle_dump (built by `make -C oracle ref` from /root/reference). Each
fixture holds, for a short seeded synthetic input, everything the reference's Predictor
produced bit by bit: Predict() outputs, the replayed PAQ8/FXCM codes and PPMD distributions,
the 47 mixer outputs and selector contexts, the small-model / PPMD / LSTM layer-0 inputs and
the LSTM byte distributions.
"""
import os, subprocess, sys, tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from gen_synth import synth_text, synth_binary  # noqa: E402
from oracle_io import Dump  # noqa: E402

CASES = [
    # name, generator, n_bytes, seed, mode
    ("text208", synth_text, 208, 0xE9E80001, "n"),
    ("binary120", synth_binary, 120, 0xE9E80003, "n"),
]


def main():
    dump_bin = os.path.join(ROOT, "oracle", "_ref", "oracle_dump")
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, gen, n, seed, mode in CASES:
        with tempfile.TemporaryDirectory() as tmp:
            src = os.path.join(tmp, "in.bin")
            open(src, "wb").write(gen(n, seed))
            subprocess.run([dump_bin, "dump", mode, src, os.path.join(tmp, "d"), "2"], check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            d = Dump(os.path.join(tmp, "d"))
            cols = list(range(3)) + list(range(2025, 2078))
            np.savez_compressed(
                os.path.join(out_dir, name + ".npz"),
                stream=d.stream, vocab=d.vocab, p=d.p, ext=d.ext, ppmd=d.ppmd, lstm=d.lstm, mix=d.mix, ctx=d.ctx,
                small_inputs=np.ascontiguousarray(d.inputs[:, cols]),
                inputs_first64=np.ascontiguousarray(d.inputs[:64]),
                generator=np.array([name, gen.__name__, str(n), hex(seed), mode]))
            print(name, "bytes", d.n_bytes, "->", os.path.getsize(os.path.join(out_dir, name + ".npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()

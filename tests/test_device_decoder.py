"""Decompress direction on the device (cmixb200_decode_bytes): the arithmetic decoder runs between the
predict and the perceive kernels of every bit, so a stream decodes without the bit visiting the host.
 * archive written by the device ENCODER (coder_begin / code_bytes / coder_finish) -> device decoder gives the bytes back;
 * the same archive through the host-side reference decoder arithmetic (tests/_coder.py restates decoder.cpp) driving the
   lock-step Predict()/Perceive() gives the same bytes: the two decode paths agree bit for bit;
 * the C-ABI exports the symbol (CPU check)."""
import os
import time

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "full_text.npz")


def test_symbol_is_exported():
    from cmix_b200.capi import load_library
    lib = load_library()
    assert hasattr(lib, "cmixb200_decode_bytes")


def host_decode(P, archive, n_bytes):
    """Decoder::Decode (reference src/coder/decoder.cpp:3-39) on the host over lock-step Predict()/Perceive()."""
    arch = bytes(archive)
    pos = 0

    def read():
        nonlocal pos
        b = arch[pos] if pos < len(arch) else 0
        pos += 1
        return b
    x1, x2, x = 0, 0xFFFFFFFF, 0
    for _ in range(4):
        x = ((x << 8) + read()) & 0xFFFFFFFF
    out = bytearray()
    for _ in range(n_bytes):
        byte = 0
        for _ in range(8):
            p = int(np.float32(1.0) + np.float32(65534.0) * np.float32(P.Predict()))
            rng = x2 - x1
            xmid = (x1 + (rng >> 16) * p + (((rng & 0xFFFF) * p) >> 16)) & 0xFFFFFFFF
            if x <= xmid:
                bit, x2 = 1, xmid
            else:
                bit, x1 = 0, (xmid + 1) & 0xFFFFFFFF
            P.Perceive(bit)
            while ((x1 ^ x2) & 0xFF000000) == 0:
                x1 = (x1 << 8) & 0xFFFFFFFF
                x2 = ((x2 << 8) + 255) & 0xFFFFFFFF
                x = ((x << 8) + read()) & 0xFFFFFFFF
            byte = byte * 2 + bit
        out.append(byte)
    return bytes(out)


@pytest.mark.gpu
@pytest.mark.timeout(900)
def test_device_decoder_round_trip():
    import cmix_b200
    os.environ.setdefault("CMIXB200_PPMD_MB", "512")
    g = np.load(GOLD)
    stream = g["stream"][:1536]
    enc = cmix_b200.Predictor(g["vocab"])
    enc.coder_begin(2 * stream.size + 64)
    p_enc = enc.code_bytes(stream)
    archive = enc.coder_finish()
    enc.close()
    np.testing.assert_array_equal(p_enc.view(np.uint32), g["p"][:stream.size * 8].view(np.uint32))   # the encoder side is the reference's

    dec = cmix_b200.Predictor(g["vocab"])
    t0 = time.perf_counter()
    out = dec.decode_bytes(archive, stream.size)
    dt = time.perf_counter() - t0
    dec.close()
    assert out.tobytes() == stream.tobytes()
    print("device decode: %.1f us per bit" % (dt / (stream.size * 8) * 1e6))

    host = cmix_b200.Predictor(g["vocab"])
    n_host = 96
    assert host_decode(host, archive, n_host) == stream[:n_host].tobytes()
    host.close()


@pytest.mark.gpu
@pytest.mark.timeout(900)
def test_device_decoder_continues_a_stream():
    """Two decode calls on one predictor, the second starting where the first stopped, equal one call; a bulk call may follow."""
    import cmix_b200
    os.environ.setdefault("CMIXB200_PPMD_MB", "512")
    g = np.load(GOLD)
    stream = g["stream"][:512]
    enc = cmix_b200.Predictor(g["vocab"])
    enc.coder_begin(2 * stream.size + 64)
    enc.code_bytes(stream)
    archive = enc.coder_finish()
    p_next = enc.code_bytes(g["stream"][512:544])
    enc.close()
    dec = cmix_b200.Predictor(g["vocab"])
    out = dec.decode_bytes(archive, stream.size)
    assert out.tobytes() == stream.tobytes()
    # the decoder's model state after 512 bytes is the encoder's: it predicts the following bytes identically
    np.testing.assert_array_equal(dec.code_bytes(g["stream"][512:544]).view(np.uint32), p_next.view(np.uint32))
    dec.close()

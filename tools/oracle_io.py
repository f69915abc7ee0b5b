"""Helpers to load oracle dumps and the CPU port (TEST INFRASTRUCTURE, used by tests/ and bench.py only)."""
import ctypes, os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_EXT, N_IN, N_MIX = 2022, 2078, 47


def load_port():
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "liboracle_port.so"))
    c = ctypes
    lib.op_create.restype = c.c_void_p
    lib.op_create.argtypes = [c.c_void_p]
    lib.op_destroy.argtypes = [c.c_void_p]
    lib.op_predict.restype = c.c_float
    lib.op_predict.argtypes = [c.c_void_p, c.c_void_p]
    lib.op_perceive.argtypes = [c.c_void_p, c.c_int, c.c_void_p]
    lib.op_pretrain.argtypes = [c.c_void_p, c.c_int]
    for n in ("op_get_inputs", "op_get_mixer_outputs", "op_get_mixer_contexts", "op_get_lstm_probs"):
        getattr(lib, n).argtypes = [c.c_void_p, c.c_void_p]
    lib.op_run.argtypes = [c.c_void_p, c.c_void_p, c.c_size_t, c.c_void_p, c.c_void_p, c.c_void_p]
    lib.op_enc_create.restype = c.c_void_p
    lib.op_enc_encode.argtypes = [c.c_void_p, c.c_float, c.c_int]
    lib.op_enc_finish.restype = c.c_size_t
    lib.op_enc_finish.argtypes = [c.c_void_p, c.c_void_p, c.c_size_t]
    lib.op_enc_destroy.argtypes = [c.c_void_p]
    lib.op_dec_create.restype = c.c_void_p
    lib.op_dec_create.argtypes = [c.c_void_p, c.c_size_t]
    lib.op_dec_decode.restype = c.c_int
    lib.op_dec_decode.argtypes = [c.c_void_p, c.c_float]
    lib.op_dec_destroy.argtypes = [c.c_void_p]
    for n in ("op_libm_expf", "op_libm_tanhf", "op_logistic"):
        getattr(lib, n).restype = c.c_float
        getattr(lib, n).argtypes = [c.c_float]
    return lib


class Dump:
    """A dump written by oracle/_ref/oracle_dump (see oracle/ref_driver.cpp)."""

    def __init__(self, prefix):
        self.prefix = prefix
        meta = {}
        for line in open(prefix + ".meta"):
            k, v = line.split(None, 1)
            meta[k] = v.strip()
        self.meta = meta
        self.n_bytes = int(meta["n_bytes"])
        self.level = int(meta["level"])
        self.vocab = np.array([int(ch) for ch in meta["vocab"]], dtype=np.uint8)
        self.stream = np.fromfile(prefix + ".stream", dtype=np.uint8)[: self.n_bytes]
        self.p = np.fromfile(prefix + ".p.f32", dtype=np.float32)
        nb = self.n_bytes * 8
        assert self.p.size == nb
        if self.level >= 1:
            self.ext = np.fromfile(prefix + ".ext.u16", dtype=np.uint16).reshape(nb, N_EXT)
            self.ppmd = np.fromfile(prefix + ".ppmd.f32", dtype=np.float32).reshape(self.n_bytes, 256)
        if self.level >= 2:
            self.inputs = np.memmap(prefix + ".in.f32", dtype=np.float32, mode="r").reshape(nb, N_IN)
            self.mix = np.fromfile(prefix + ".mix.f32", dtype=np.float32).reshape(nb, N_MIX)
            self.ctx = np.fromfile(prefix + ".ctx.u32", dtype=np.uint32).reshape(nb, N_MIX)
            self.lstm = np.fromfile(prefix + ".lstm.f32", dtype=np.float32).reshape(self.n_bytes, 256)

    def bits(self):
        return np.unpackbits(self.stream)

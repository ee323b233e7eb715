"""ctypes bindings for the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

  * ``Port``  -> oracle/libria_oracle.so  (our plain-C restatement, oracle/*.c)
  * ``Ref``   -> oracle/_ref/libria_ref.so (the UNMODIFIED reference + oracle/ref_shim.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package ria_b200 never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libria_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libria_ref.so")

LDPC_N = 648
# CodeRate numeric values, include/ultra/types.hpp:91-100
R1_4, R1_3, R1_2, R2_3, R3_4, R5_6, R7_8 = range(7)
RATE_K = {R1_4: 162, R1_2: 324, R2_3: 432, R3_4: 486, R5_6: 540, R1_3: 324, R7_8: 324}
# LDPCCodec::getRecommendedIterations, src/fec/ldpc_codec.hpp:86-96
RATE_MAX_ITER = {R1_4: 50, R1_3: 60, R1_2: 80, R2_3: 70, R3_4: 60, R5_6: 50, R7_8: 50}

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build(ref: bool = True) -> None:
    """Build the checkers (make -C oracle).  Building the checker is not using it."""
    targets = ["port"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", HERE, "-j8"] + targets, check=True)


class _Code(C.Structure):
    _fields_ = [("rate", C.c_int), ("k", C.c_int), ("m", C.c_int), ("n", C.c_int),
                ("n_edges", C.c_int), ("row_ptr", C.c_int * (LDPC_N + 1)),
                ("edge_var", C.c_int * 4096)]


class Port:
    """Plain-C restatement (oracle/*.c)."""

    def __init__(self):
        if not os.path.exists(PORT_SO):
            build(ref=False)
        self.lib = L = C.CDLL(PORT_SO)
        L.orc_ldpc_build.argtypes = [C.c_int, C.POINTER(_Code)]
        L.orc_ldpc_encode.argtypes = [C.POINTER(_Code), _u8p, C.c_int, _u8p, C.c_int]
        L.orc_ldpc_encode.restype = C.c_int
        L.orc_ldpc_decode_batch.argtypes = [C.POINTER(_Code), _f32p, C.c_int, C.c_int, C.c_float,
                                            _u8p, C.c_int, _u8p, _i32p]
        L.orc_crc16.argtypes = [_u8p, C.c_int]
        L.orc_crc16.restype = C.c_uint16
        L.orc_mt_seed.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_mt_next.argtypes = [C.c_void_p]
        L.orc_mt_next.restype = C.c_uint32
        self._codes = {}

    def code(self, rate: int) -> _Code:
        if rate not in self._codes:
            c = _Code()
            self.lib.orc_ldpc_build(rate, C.byref(c))
            self._codes[rate] = c
        return self._codes[rate]

    def ldpc_edges(self, rate: int):
        c = self.code(rate)
        row_ptr = np.array(c.row_ptr[: c.m + 1], dtype=np.int32)
        edge_var = np.array(c.edge_var[: c.n_edges], dtype=np.int32)
        return c.k, c.m, row_ptr, edge_var

    def ldpc_encode(self, rate: int, data: np.ndarray) -> np.ndarray:
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.zeros(((len(data) * 8 + 161) // 162 + 1) * 81, dtype=np.uint8)
        n = self.lib.orc_ldpc_encode(C.byref(self.code(rate)), data, len(data), out, len(out))
        assert n >= 0
        return out[:n].copy()

    def ldpc_decode_batch(self, rate: int, llr: np.ndarray, max_iter: int, factor: float,
                          out_stride: int = 64):
        llr = np.ascontiguousarray(llr, dtype=np.float32).reshape(-1, LDPC_N)
        n = llr.shape[0]
        out = np.zeros((n, out_stride), dtype=np.uint8)
        ok = np.zeros(n, dtype=np.uint8)
        iters = np.zeros(n, dtype=np.int32)
        self.lib.orc_ldpc_decode_batch(C.byref(self.code(rate)), llr, n, max_iter, factor,
                                       out, out_stride, ok, iters)
        return out, ok, iters

    def crc16(self, data) -> int:
        data = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8))
        return int(self.lib.orc_crc16(data, len(data)))


class Ref:
    """The unmodified reference behind oracle/ref_shim.cpp."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            build(ref=True)
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO)
        self.lib = L = C.CDLL(REF_SO)
        L.ref_quiet()
        L.ref_ldpc_encode.argtypes = [C.c_int, _u8p, C.c_int, _u8p, C.c_int]
        L.ref_ldpc_encode.restype = C.c_int
        L.ref_ldpc_decoder_new.argtypes = [C.c_int, C.c_int, C.c_float]
        L.ref_ldpc_decoder_new.restype = C.c_void_p
        L.ref_ldpc_decoder_free.argtypes = [C.c_void_p]
        L.ref_ldpc_decode_soft.argtypes = [C.c_void_p, _f32p, C.c_int, _u8p, C.c_int,
                                           C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ref_ldpc_decode_soft.restype = C.c_int
        L.ref_ldpc_decode_batch.argtypes = [C.c_void_p, _f32p, C.c_int, _u8p, C.c_int, _u8p, _i32p]

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_SO)

    def ldpc_encode(self, rate: int, data: np.ndarray) -> np.ndarray:
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.zeros(((len(data) * 8 + 161) // 162 + 1) * 81, dtype=np.uint8)
        n = self.lib.ref_ldpc_encode(rate, data, len(data), out, len(out))
        assert n >= 0
        return out[:n].copy()

    def ldpc_decode_soft(self, rate: int, llr: np.ndarray, max_iter: int, factor: float):
        llr = np.ascontiguousarray(llr, dtype=np.float32).ravel()
        h = self.lib.ref_ldpc_decoder_new(rate, max_iter, factor)
        out = np.zeros(len(llr) // 8 + 82, dtype=np.uint8)
        ok, it = C.c_int(0), C.c_int(0)
        n = self.lib.ref_ldpc_decode_soft(h, llr, len(llr), out, len(out), C.byref(ok), C.byref(it))
        self.lib.ref_ldpc_decoder_free(h)
        return out[:n].copy(), bool(ok.value), it.value

    def ldpc_decode_batch(self, rate: int, llr: np.ndarray, max_iter: int, factor: float,
                          out_stride: int = 64):
        llr = np.ascontiguousarray(llr, dtype=np.float32).reshape(-1, LDPC_N)
        n = llr.shape[0]
        out = np.zeros((n, out_stride), dtype=np.uint8)
        ok = np.zeros(n, dtype=np.uint8)
        iters = np.zeros(n, dtype=np.int32)
        h = self.lib.ref_ldpc_decoder_new(rate, max_iter, factor)
        self.lib.ref_ldpc_decode_batch(h, llr, n, out, out_stride, ok, iters)
        self.lib.ref_ldpc_decoder_free(h)
        return out, ok, iters


# ---------------------------------------------------------------------------------------------
# Synthetic inputs (shared by tests and bench; numpy only, deterministic)
# ---------------------------------------------------------------------------------------------

def unpack_bits(coded: np.ndarray, n: int = LDPC_N) -> np.ndarray:
    return np.unpackbits(np.asarray(coded, dtype=np.uint8))[:n]


def awgn_llrs(bits: np.ndarray, esn0_db: float, rng: np.random.Generator) -> np.ndarray:
    """LLR model of tools/test_chase_cache.cpp:20-34: llr = 2*(s+n)*snr, n ~ N(0, 1/snr)."""
    snr = np.float32(10.0 ** (esn0_db / 10.0))
    s = 1.0 - 2.0 * bits.astype(np.float32)
    noise = rng.standard_normal(bits.shape, dtype=np.float32) / np.sqrt(snr)
    return (2.0 * (s + noise) * snr).astype(np.float32)

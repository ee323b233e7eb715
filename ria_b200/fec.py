"""Host-side mirror of the reference FEC interfaces over the C ABI.

Same names, argument meaning and error behaviour as the reference classes so parity tests read
like the reference's own:

  * ``LDPCDecoder``  <- ultra::LDPCDecoder (include/ultra/fec.hpp:48-81,
                        src/fec/ldpc_decoder.cpp:263-455)
  * ``LDPCCodec``    <- ultra::fec::LDPCCodec / ICodec (src/fec/ldpc_codec.hpp:38-105,
                        src/fec/ldpc_codec.cpp:47-129, src/fec/codec_interface.hpp:29-98)

plus the batched entry point the GPU exists for (``decode_batch``).  Device memory and streams
come from torch; all arithmetic happens in libria_b200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np
import torch

from ._lib import Context, RiaError, lib

LDPC_N = 648

# ultra::CodeRate (include/ultra/types.hpp:91-100)
R1_4, R1_3, R1_2, R2_3, R3_4, R5_6, R7_8 = range(7)


def code_params(rate: int) -> Tuple[int, int, int]:
    k, m, e = C.c_int(), C.c_int(), C.c_int()
    if lib().ria_ldpc_params(int(rate), C.byref(k), C.byref(m), C.byref(e)) != 0:
        raise ValueError(f"bad code rate {rate}")
    return k.value, m.value, e.value


def get_matrix(rate: int):
    """(row_ptr[m+1], edge_var[E]) of the H the library generated (H_rows order)."""
    k, m, e = code_params(rate)
    row_ptr = np.zeros(m + 1, dtype=np.int32)
    edge_var = np.zeros(e, dtype=np.int32)
    rc = lib().ria_ldpc_get_matrix(int(rate), row_ptr.ctypes.data, edge_var.ctypes.data)
    if rc != 0:
        raise ValueError(f"bad code rate {rate}")
    return row_ptr, edge_var


def recommended_iterations(rate: int) -> int:
    """LDPCCodec::getRecommendedIterations (src/fec/ldpc_codec.hpp:86-96)."""
    return {R3_4: 60, R2_3: 70, R1_2: 80, R1_3: 60, R1_4: 50}.get(int(rate), 50)


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        if not torch.cuda.is_available():
            raise RiaError("no CUDA device: ria_b200 has no CPU fallback")
        _default_ctx = Context(torch.cuda.current_device())
    return _default_ctx


class LDPCDecoder:
    """ultra::LDPCDecoder drop-in (batch = 1 semantics) + ``decode_batch``."""

    def __init__(self, rate: int, ctx: Optional[Context] = None):
        code_params(rate)
        self._rate = int(rate)
        self._ctx = ctx
        self._max_iter = 50          # Impl::max_iterations default, ldpc_decoder.cpp:43
        self._factor = 0.75          # Impl::min_sum_factor default, ldpc_decoder.cpp:44
        self._last_success = False
        self._last_iters = 0

    # ---- reference setters / getters (ldpc_decoder.cpp:431-455) ----
    def setRate(self, rate: int) -> None:
        code_params(rate)
        self._rate = int(rate)

    def getRate(self) -> int:
        return self._rate

    def setMaxIterations(self, max_iter: int) -> None:
        self._max_iter = int(max_iter)

    def setMinSumFactor(self, factor: float) -> None:
        self._factor = float(factor)

    def lastDecodeSuccess(self) -> bool:
        return self._last_success

    def lastIterations(self) -> int:
        return self._last_iters

    @property
    def ctx(self) -> Context:
        if self._ctx is None:
            self._ctx = default_context()
        return self._ctx

    # ---- batched device entry point ----
    def decode_batch(self, llr: torch.Tensor, info_stride: Optional[int] = None):
        """llr: CUDA fp32 [n_cw, 648] -> (info u8 [n_cw, stride], ok u8 [n_cw], iters i32 [n_cw])."""
        if not (isinstance(llr, torch.Tensor) and llr.is_cuda):
            raise RiaError("decode_batch wants a CUDA tensor (no CPU fallback); use decode_batch_host "
                           "for host buffers")
        if llr.dtype != torch.float32 or llr.dim() != 2 or llr.shape[1] != LDPC_N:
            raise ValueError("llr must be fp32 [n_cw, 648]")
        llr = llr.contiguous()
        k, _, _ = code_params(self._rate)
        stride = int(info_stride) if info_stride else (k + 7) // 8
        n = llr.shape[0]
        info = torch.empty((n, stride), dtype=torch.uint8, device=llr.device)
        ok = torch.empty((n,), dtype=torch.uint8, device=llr.device)
        iters = torch.empty((n,), dtype=torch.int32, device=llr.device)
        ctx = self.ctx
        ctx.set_stream(torch.cuda.current_stream(llr.device))
        ctx.check(lib().ria_ldpc_decode_batch_dev(
            ctx.handle, self._rate, self._max_iter, self._factor, llr.data_ptr(), n,
            info.data_ptr(), stride, ok.data_ptr(), iters.data_ptr()))
        return info, ok, iters

    def decode_batch_host(self, llr: np.ndarray, info_stride: Optional[int] = None):
        """Host buffers in, host buffers out (H2D/D2H inside the C call)."""
        llr = np.ascontiguousarray(llr, dtype=np.float32).reshape(-1, LDPC_N)
        k, _, _ = code_params(self._rate)
        stride = int(info_stride) if info_stride else (k + 7) // 8
        n = llr.shape[0]
        info = np.empty((n, stride), dtype=np.uint8)
        ok = np.empty((n,), dtype=np.uint8)
        iters = np.empty((n,), dtype=np.int32)
        ctx = self.ctx
        ctx.check(lib().ria_ldpc_decode_batch_host(
            ctx.handle, self._rate, self._max_iter, self._factor, llr.ctypes.data, n,
            info.ctypes.data, stride, ok.ctypes.data, iters.ctypes.data))
        return info, ok, iters

    # ---- reference single-call semantics (ldpc_decoder.cpp:284-429) ----
    def decodeSoft(self, llrs) -> bytes:
        llrs = np.asarray(llrs, dtype=np.float32).ravel()
        if llrs.size == 0:                       # :286-289
            self._last_success = False
            return b""
        k, _, _ = code_params(self._rate)
        n_full = llrs.size // LDPC_N
        rem = llrs.size - n_full * LDPC_N
        single = llrs.size <= LDPC_N             # :297-300
        n_blocks = 1 if single else n_full + (1 if rem else 0)
        padded = np.zeros((n_blocks, LDPC_N), dtype=np.float32)
        padded.ravel()[: llrs.size] = llrs
        info, ok, iters = self.decode_batch_host(padded)
        self._last_iters = int(iters[-1])
        if single:
            self._last_success = bool(ok[0])
            return info[0, : (k + 7) // 8].tobytes()
        # multi-block: bit-level concatenation of the k info bits of each block (:302-428).
        # last_success: AND over whole blocks, but a trailing partial block goes through
        # decodeBP which overwrites the flag with its own result (:401).
        self._last_success = bool(ok[-1]) if rem else bool(ok[:n_full].all())
        bits = np.unpackbits(info[:, : (k + 7) // 8], axis=1)[:, :k].ravel()
        return np.packbits(bits).tobytes()

    def decode(self, coded: bytes) -> bytes:
        """Hard-bit entry (ldpc_decoder.cpp:268-282): bit -> -6/+6 LLR."""
        bits = np.unpackbits(np.frombuffer(bytes(coded), dtype=np.uint8))
        return self.decodeSoft(np.where(bits == 1, -6.0, 6.0).astype(np.float32))


@dataclass
class DecodeResult:
    """fec::DecodeResult (src/fec/codec_interface.hpp:29-34)."""
    success: bool = False
    data: bytes = b""
    iterations: int = 0
    ber_estimate: float = 0.0


class LDPCCodec:
    """fec::LDPCCodec / ICodec decode side (src/fec/ldpc_codec.cpp:47-129)."""

    CODEWORD_BITS = 648
    CODEWORD_BYTES = 81

    def __init__(self, rate: int = R1_2, ctx: Optional[Context] = None):
        self._rate = int(rate)
        self._max_iterations = recommended_iterations(rate)
        self._decoder = LDPCDecoder(rate, ctx)
        self._decoder.setMaxIterations(self._max_iterations)

    def getName(self) -> str:
        return "802.11n LDPC"

    def setRate(self, rate: int) -> None:
        self._rate = int(rate)
        self._decoder.setRate(rate)
        rec = recommended_iterations(rate)
        if self._max_iterations != rec:
            self._max_iterations = rec
            self._decoder.setMaxIterations(rec)

    def getRate(self) -> int:
        return self._rate

    def setMaxIterations(self, iterations: int) -> None:
        self._max_iterations = int(iterations)
        self._decoder.setMaxIterations(iterations)

    def getMaxIterations(self) -> int:
        return self._max_iterations

    def decode(self, soft_bits) -> Tuple[bool, bytes]:
        data = self._decoder.decodeSoft(soft_bits)
        return self._decoder.lastDecodeSuccess(), data

    def decodeExtended(self, soft_bits) -> DecodeResult:
        r = DecodeResult()
        r.data = self._decoder.decodeSoft(soft_bits)
        r.success = self._decoder.lastDecodeSuccess()
        r.iterations = self._decoder.lastIterations()
        r.ber_estimate = (np.float32(r.iterations) / np.float32(self._max_iterations * 10.0)
                          if r.success else 0.5)
        return r

    def getCodewordBits(self) -> int:
        return self.CODEWORD_BITS

    def getInfoBits(self) -> int:
        return code_params(self._rate)[0]

    def getParityBits(self) -> int:
        return self.CODEWORD_BITS - self.getInfoBits()

    def getCodewordBytes(self) -> int:
        return self.CODEWORD_BYTES

    def getDataBytes(self) -> int:
        return self.getInfoBits() // 8

    def getEffectiveRate(self) -> float:
        return self.getInfoBits() / self.CODEWORD_BITS

    @property
    def decoder(self) -> LDPCDecoder:
        return self._decoder

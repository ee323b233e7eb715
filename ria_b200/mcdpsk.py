"""Host-side mirror of the reference MC-DPSK receive interface over the C ABI.

  * ``MultiCarrierDPSKConfig``  <- ultra::MultiCarrierDPSKConfig (src/psk/multi_carrier_dpsk.hpp:27-100)
  * ``MCDPSKDemodulator``       <- MCDPSKWaveform::process / MultiCarrierDPSKDemodulator
                                   (src/waveform/mc_dpsk_waveform.cpp:294-338,
                                    src/psk/multi_carrier_dpsk.hpp:797-895), batched.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from ._lib import Context, RiaError, lib
from .fec import default_context

SPREAD_NONE, SPREAD_2X, SPREAD_4X = 1, 2, 4


class MultiCarrierDPSKConfig(C.Structure):
    """ria_mcdpsk_config (include/ria_b200.h)."""
    _fields_ = [("sample_rate", C.c_float), ("num_carriers", C.c_uint32), ("freq_low", C.c_float),
                ("freq_high", C.c_float), ("samples_per_symbol", C.c_uint32),
                ("bits_per_symbol", C.c_uint32), ("spreading", C.c_uint32),
                ("training_symbols", C.c_uint32)]

    @classmethod
    def default(cls, bits_per_symbol=2, spreading=SPREAD_NONE, num_carriers=8, training_symbols=8):
        return cls(48000.0, num_carriers, 500.0, 2500.0, 512, bits_per_symbol, spreading, training_symbols)

    @classmethod
    def level4_dbpsk(cls, spreading=SPREAD_NONE):
        """mc_dpsk_presets::level4_dbpsk: 10 carriers DBPSK (multi_carrier_dpsk.hpp:996-1003)."""
        return cls.default(1, spreading, 10)

    def soft_bits_per_frame(self, frame_len: int) -> int:
        n = lib().ria_mcdpsk_soft_bits_per_frame(C.addressof(self), int(frame_len))
        if n < 0:
            raise ValueError("unsupported MC-DPSK configuration")
        return n

    def preamble_samples(self) -> int:
        return (self.training_symbols + 1) * self.samples_per_symbol


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class MCDPSKDemodulator:
    def __init__(self, config: MultiCarrierDPSKConfig, ctx: Optional[Context] = None):
        self.config = config
        self._ctx = ctx
        config.soft_bits_per_frame(0)

    @property
    def ctx(self) -> Context:
        if self._ctx is None:
            self._ctx = default_context()
        return self._ctx

    def process_batch(self, samples: torch.Tensor, cfo_hz: Optional[torch.Tensor] = None,
                      phase: Optional[torch.Tensor] = None):
        """samples: CUDA fp32 [n_frames, frame_len] = [training][reference][data] per row.

        Returns dict(llr [n, stride], n_llr [n], fading [n], cfo [n])."""
        if not (isinstance(samples, torch.Tensor) and samples.is_cuda and samples.dtype == torch.float32
                and samples.dim() == 2):
            raise RiaError("process_batch wants CUDA fp32 [n_frames, frame_len] (no CPU fallback)")
        if samples.stride(1) != 1:
            samples = samples.contiguous()
        n, frame_len = samples.shape
        dev = samples.device
        n_llr = self.config.soft_bits_per_frame(frame_len)
        stride = max(4, (n_llr + 3) & ~3)
        out = dict(llr=torch.empty((n, stride), dtype=torch.float32, device=dev),
                   n_llr=torch.empty((n,), dtype=torch.int32, device=dev),
                   fading=torch.empty((n,), dtype=torch.float32, device=dev),
                   cfo=torch.empty((n,), dtype=torch.float32, device=dev))
        ctx = self.ctx
        ctx.set_stream(torch.cuda.current_stream(dev))
        ctx.check(lib().ria_mcdpsk_process_batch_dev(
            ctx.handle, C.addressof(self.config), _ptr(samples), samples.stride(0), frame_len,
            _ptr(cfo_hz), _ptr(phase), n, _ptr(out["llr"]), stride, _ptr(out["n_llr"]),
            _ptr(out["fading"]), _ptr(out["cfo"])))
        return out


class McdpskRxChain:
    """Chirp-acquired MC-DPSK frames, one codeword per frame (BASELINE configs[2]):
    MCDPSKWaveform::detectSync -> process -> ChaseCache combining -> LDPCDecoder::decodeSoft,
    all on the device through ``ria_mcdpsk_rx_frames_dev`` / ``_host``."""

    def __init__(self, config: MultiCarrierDPSKConfig, rate: int, max_iter: int = 50, min_sum_factor: float = 0.9375,
                 threshold: float = 0.15, ctx: Optional[Context] = None):
        from .sync import ChirpConfig
        self.config = config
        self.chirp = ChirpConfig.default()
        self.rate, self.max_iter, self.factor, self.threshold = int(rate), int(max_iter), float(min_sum_factor), float(threshold)
        self._ctx = ctx
        from . import fec
        self.info_bytes = (fec.code_params(self.rate)[0] + 7) // 8

    @property
    def ctx(self) -> Context:
        if self._ctx is None:
            self._ctx = default_context()
        return self._ctx

    def process_batch(self, rows: torch.Tensor, frame_len: int, sync_window: int, acc: Optional[torch.Tensor] = None,
                      first_reception: bool = True, out=None):
        """rows: CUDA fp32 [n, row_len].  acc: CUDA fp32 [n, 648] chase accumulators (created when None).

        Returns dict(info u8 [n, info_stride], ok u8 [n], iters i32 [n], sync u8 [n, 32], acc)."""
        from .sync import SYNC_RESULT_DTYPE
        if not (isinstance(rows, torch.Tensor) and rows.is_cuda and rows.dtype == torch.float32 and rows.dim() == 2):
            raise RiaError("process_batch wants CUDA fp32 [n_frames, row_len] (no CPU fallback)")
        if rows.stride(1) != 1:
            rows = rows.contiguous()
        n = rows.shape[0]
        dev = rows.device
        info_stride = (self.info_bytes + 3) & ~3
        if acc is None:
            acc = torch.empty((n, 648), dtype=torch.float32, device=dev)
        if out is None:
            out = dict(info=torch.empty((n, info_stride), dtype=torch.uint8, device=dev),
                       ok=torch.empty((n,), dtype=torch.uint8, device=dev),
                       iters=torch.empty((n,), dtype=torch.int32, device=dev),
                       sync=torch.empty((n, SYNC_RESULT_DTYPE.itemsize), dtype=torch.uint8, device=dev))
        out["acc"] = acc
        ctx = self.ctx
        ctx.set_stream(torch.cuda.current_stream(dev))
        ctx.check(lib().ria_mcdpsk_rx_frames_dev(
            ctx.handle, C.addressof(self.config), C.addressof(self.chirp), _ptr(rows), rows.stride(0),
            int(sync_window), int(frame_len), self.threshold, n, self.rate, self.max_iter, self.factor,
            _ptr(acc), int(bool(first_reception)), _ptr(out["info"]), out["info"].stride(0), _ptr(out["ok"]),
            _ptr(out["iters"]), _ptr(out["sync"])))
        return out

    def process_batch_host(self, rows, frame_len: int, sync_window: int):
        """rows: host fp32 array or (pinned) CPU tensor [n, row_len]; H2D, chain and D2H happen inside the call."""
        import numpy as np
        from .sync import SYNC_RESULT_DTYPE
        t = rows if isinstance(rows, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(rows, dtype=np.float32))
        if t.is_cuda or t.dtype != torch.float32 or t.dim() != 2 or t.stride(1) != 1:
            raise RiaError("process_batch_host wants a contiguous host fp32 [n_frames, row_len]")
        n = t.shape[0]
        info_stride = (self.info_bytes + 3) & ~3
        info = np.empty((n, info_stride), np.uint8)
        ok = np.empty(n, np.uint8)
        iters = np.empty(n, np.int32)
        sync = np.empty(n, SYNC_RESULT_DTYPE)
        ctx = self.ctx
        ctx.check(lib().ria_mcdpsk_rx_frames_host(
            ctx.handle, C.addressof(self.config), C.addressof(self.chirp), C.c_void_p(t.data_ptr()), t.stride(0),
            int(sync_window), int(frame_len), self.threshold, n, self.rate, self.max_iter, self.factor,
            info.ctypes.data_as(C.c_void_p), info_stride, ok.ctypes.data_as(C.c_void_p),
            iters.ctypes.data_as(C.c_void_p), sync.ctypes.data_as(C.c_void_p)))
        return dict(info=info, ok=ok, iters=iters, sync=sync)


def mcdpsk_tx_frames(config: "MultiCarrierDPSKConfig", data: torch.Tensor, ctx: Optional[Context] = None) -> torch.Tensor:
    """MultiCarrierDPSKModulator training + reference + modulate(data) for a batch on the device
    (src/psk/multi_carrier_dpsk.hpp:141-275): data CUDA u8 [n, len] -> samples fp32 [n, frame_len],
    sample-identical to the reference transmitter."""
    if not (isinstance(data, torch.Tensor) and data.is_cuda and data.dtype == torch.uint8 and data.dim() == 2):
        raise RiaError("mcdpsk_tx_frames wants a CUDA uint8 [n, len] tensor")
    data = data.contiguous()
    ctx = ctx or default_context()
    n, ln = data.shape
    flen = lib().ria_mcdpsk_tx_frame_samples(C.addressof(config), ln)
    if flen <= 0:
        raise RiaError("mcdpsk_tx_frames: unsupported configuration")
    out = torch.empty((n, flen), dtype=torch.float32, device=data.device)
    ctx.set_stream(torch.cuda.current_stream(data.device))
    ctx.check(lib().ria_mcdpsk_tx_frames_dev(ctx.handle, C.addressof(config), data.data_ptr(), data.stride(0), ln, n,
                                             out.data_ptr(), out.stride(0)))
    return out

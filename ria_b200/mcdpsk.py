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

"""Host-side mirror of the reference synchronisers over the C ABI.

  * ``ZCConfig`` / ``ZCSync.detect_batch``      <- sync::ZCSync::detect (src/sync/zc_sync.hpp:192-391)
  * ``ChirpSync.detect_dual_batch``             <- sync::ChirpSync::detectDualChirp
                                                   (src/sync/chirp_sync.hpp:352-512)
Results come back as a structured numpy array with the fields of ``ria_sync_result``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from ._lib import Context, RiaError, lib
from .fec import default_context

ZC_ROOT_MASK_PING, ZC_ROOT_MASK_PONG, ZC_ROOT_MASK_DATA, ZC_ROOT_MASK_CONTROL = 1, 2, 4, 8
ZC_ROOT_MASK_ALL = 15
ZC_DEFAULT_DETECT_THRESHOLD = 0.3

SYNC_RESULT_DTYPE = np.dtype([("detected", np.int32), ("start_sample", np.int32), ("correlation", np.float32),
                              ("cfo_hz", np.float32), ("snr_estimate", np.float32), ("root", np.int32),
                              ("frame_type", np.int32), ("aux", np.int32)])


class ZCConfig(C.Structure):
    """ria_zc_config (include/ria_b200.h) = sync::ZCConfig."""
    _fields_ = [("sample_rate", C.c_float), ("sequence_length", C.c_int32), ("upsample_factor", C.c_int32),
                ("num_repetitions", C.c_int32), ("carrier_freq", C.c_float), ("gap_ms", C.c_float),
                ("root_ping", C.c_int32), ("root_pong", C.c_int32), ("root_data", C.c_int32),
                ("root_control", C.c_int32)]

    @classmethod
    def default(cls):
        c = cls()
        lib().ria_zc_config_default(C.addressof(c))
        return c

    def singleRepSamples(self) -> int:
        return self.sequence_length * self.upsample_factor

    def preambleSamples(self) -> int:
        return self.singleRepSamples() * self.num_repetitions + int(np.float32(self.sample_rate) * np.float32(self.gap_ms) / np.float32(1000.0))


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _check_windows(samples):
    if not (isinstance(samples, torch.Tensor) and samples.is_cuda and samples.dtype == torch.float32
            and samples.dim() == 2):
        raise RiaError("detect wants CUDA fp32 [n_windows, window] (no CPU fallback)")
    return samples if samples.stride(1) == 1 else samples.contiguous()


class ZCSync:
    def __init__(self, config: Optional[ZCConfig] = None, ctx: Optional[Context] = None):
        self.config = config or ZCConfig.default()
        self._ctx = ctx

    @property
    def ctx(self) -> Context:
        if self._ctx is None:
            self._ctx = default_context()
        return self._ctx

    def detect_batch(self, samples: torch.Tensor, threshold: float = ZC_DEFAULT_DETECT_THRESHOLD,
                     root_mask: int = ZC_ROOT_MASK_ALL, known_cfo_hz: Optional[torch.Tensor] = None) -> torch.Tensor:
        """-> uint8 tensor [n, 32]; ``results(t)`` views it with SYNC_RESULT_DTYPE."""
        samples = _check_windows(samples)
        n, window = samples.shape
        out = torch.empty((n, SYNC_RESULT_DTYPE.itemsize), dtype=torch.uint8, device=samples.device)
        ctx = self.ctx
        ctx.set_stream(torch.cuda.current_stream(samples.device))
        for off in range(0, n, 32768):
            m = min(32768, n - off)
            ctx.check(lib().ria_zc_detect_batch_dev(
                ctx.handle, C.addressof(self.config), _ptr(samples[off:]), samples.stride(0), window,
                _ptr(known_cfo_hz[off:]) if known_cfo_hz is not None else C.c_void_p(0),
                float(threshold), int(root_mask), m, _ptr(out[off:])))
        return out


def results(t: torch.Tensor) -> np.ndarray:
    return t.cpu().numpy().view(SYNC_RESULT_DTYPE).reshape(-1)


class ChirpConfig(C.Structure):
    """ria_chirp_config (include/ria_b200.h) = sync::ChirpConfig with dual chirp."""
    _fields_ = [("sample_rate", C.c_float), ("f_start", C.c_float), ("f_end", C.c_float),
                ("duration_ms", C.c_float), ("gap_ms", C.c_float)]

    @classmethod
    def default(cls):
        c = cls()
        lib().ria_chirp_config_default(C.addressof(c))
        return c

    def getChirpSamples(self) -> int:
        return int(np.float32(self.sample_rate) * np.float32(self.duration_ms) / np.float32(1000.0))

    def getTotalSamples(self) -> int:
        gap = int(np.float32(self.sample_rate) * np.float32(self.gap_ms) / np.float32(1000.0))
        return 2 * self.getChirpSamples() + 2 * gap


class ChirpSync:
    def __init__(self, config: Optional[ChirpConfig] = None, ctx: Optional[Context] = None):
        self.config = config or ChirpConfig.default()
        self._ctx = ctx

    @property
    def ctx(self) -> Context:
        if self._ctx is None:
            self._ctx = default_context()
        return self._ctx

    def detect_dual_batch(self, samples: torch.Tensor, threshold: float = 0.15, max_batch: int = 2048) -> torch.Tensor:
        """detectDualChirp for every row of samples (CUDA fp32 [n, window <= 131072]).

        Each window needs 3 MiB of scratch for the 131072-point spectra, so the batch is walked
        in slices of ``max_batch`` windows."""
        samples = _check_windows(samples)
        n, window = samples.shape
        out = torch.empty((n, SYNC_RESULT_DTYPE.itemsize), dtype=torch.uint8, device=samples.device)
        ctx = self.ctx
        ctx.set_stream(torch.cuda.current_stream(samples.device))
        for off in range(0, n, max_batch):
            m = min(max_batch, n - off)
            ctx.check(lib().ria_chirp_detect_dual_batch_dev(
                ctx.handle, C.addressof(self.config), _ptr(samples[off:]), samples.stride(0), window,
                float(threshold), m, _ptr(out[off:])))
        return out


def ofdm_data_sync_batch(config, samples: torch.Tensor, known_cfo_hz: Optional[torch.Tensor] = None,
                         threshold: float = 0.3, ctx: Optional[Context] = None) -> torch.Tensor:
    """OFDMChirpWaveform::detectDataSync for every row of samples (CUDA fp32 [n, window]);
    ``config`` is a ria_b200.ofdm.ModemConfig."""
    samples = _check_windows(samples)
    n, window = samples.shape
    out = torch.empty((n, SYNC_RESULT_DTYPE.itemsize), dtype=torch.uint8, device=samples.device)
    ctx = ctx or default_context()
    ctx.set_stream(torch.cuda.current_stream(samples.device))
    ctx.check(lib().ria_ofdm_data_sync_batch_dev(
        ctx.handle, C.addressof(config), _ptr(samples), samples.stride(0), window,
        _ptr(known_cfo_hz), float(threshold), n, _ptr(out)))
    return out


def ofdm_cox_search_sync_batch(config, samples: torch.Tensor, threshold: float = 0.8,
                               noise_floor: Optional[torch.Tensor] = None, ctx: Optional[Context] = None) -> torch.Tensor:
    """OFDMDemodulator::searchForSync (src/ofdm/demodulator.cpp:1450-1542) = OFDMNvisWaveform::detectSync for every
    row of samples (CUDA fp32 [n, window <= 65536]).  Result rows: detected, start_sample = first LTS sample,
    cfo_hz, aux = Schmidl-Cox peak.  ``noise_floor`` (CUDA fp32 [n], updated in place) carries the demodulator's
    noise-floor tracker from one call to the next; None = fresh demodulators."""
    samples = _check_windows(samples)
    n, window = samples.shape
    if noise_floor is not None and not (noise_floor.is_cuda and noise_floor.dtype == torch.float32
                                        and noise_floor.is_contiguous() and noise_floor.numel() == n):
        raise RiaError("noise_floor must be a contiguous CUDA fp32 [n] tensor")
    out = torch.empty((n, SYNC_RESULT_DTYPE.itemsize), dtype=torch.uint8, device=samples.device)
    ctx = ctx or default_context()
    ctx.set_stream(torch.cuda.current_stream(samples.device))
    ctx.check(lib().ria_ofdm_cox_search_sync_batch_dev(
        ctx.handle, C.addressof(config), _ptr(samples), samples.stride(0), window, float(threshold),
        _ptr(noise_floor), n, _ptr(out)))
    return out


def ofdm_cox_correlation_batch(config, samples: torch.Tensor, offsets: torch.Tensor,
                               ctx: Optional[Context] = None) -> torch.Tensor:
    """Impl::measureCorrelation(offset) (src/ofdm/ofdm_sync.cpp:118-190), one offset (CUDA int32 [n]) per window."""
    samples = _check_windows(samples)
    n, window = samples.shape
    offsets = offsets.to(torch.int32).contiguous()
    out = torch.empty(n, dtype=torch.float32, device=samples.device)
    ctx = ctx or default_context()
    ctx.set_stream(torch.cuda.current_stream(samples.device))
    ctx.check(lib().ria_ofdm_cox_correlation_batch_dev(
        ctx.handle, C.addressof(config), _ptr(samples), samples.stride(0), window, _ptr(offsets), n, _ptr(out)))
    return out


def zc_preamble_host(config: Optional[ZCConfig] = None, root: int = 5) -> np.ndarray:
    """sync::ZCSync::generatePreambleForRoot (src/sync/zc_sync.hpp:133-190), the reference's samples
    (host evaluation; root 5 = DATA frames)."""
    config = config or ZCConfig.default()
    n = -lib().ria_zc_preamble_host(C.addressof(config), int(root), None, 0)
    out = np.empty(n, np.float32)
    got = lib().ria_zc_preamble_host(C.addressof(config), int(root), out.ctypes.data, n)
    assert got == n
    return out


def chirp_generate_host(config: Optional[ChirpConfig] = None) -> np.ndarray:
    """sync::ChirpSync::generate (src/sync/chirp_sync.hpp:61-108): [up chirp][gap][down chirp][gap], the
    reference's samples (host evaluation)."""
    config = config or ChirpConfig.default()
    n = -lib().ria_chirp_generate_host(C.addressof(config), None, 0)
    out = np.empty(n, np.float32)
    got = lib().ria_chirp_generate_host(C.addressof(config), out.ctypes.data, n)
    assert got == n
    return out


def zc_preamble(config: Optional[ZCConfig] = None, root: int = 5, device=None,
                ctx: Optional[Context] = None) -> torch.Tensor:
    """ZCSync::generatePreambleForRoot on the device (one thread per sample), CUDA fp32 [n]."""
    config = config or ZCConfig.default()
    n = lib().ria_zc_preamble_samples(C.addressof(config))
    if n <= 0:
        raise RiaError("bad ZC configuration")
    ctx = ctx or default_context()
    out = torch.empty(n, dtype=torch.float32, device=device or torch.device("cuda", ctx.device))
    ctx.set_stream(torch.cuda.current_stream(out.device))
    ctx.check(min(0, lib().ria_zc_preamble_dev(ctx.handle, C.addressof(config), int(root), _ptr(out), n)))
    return out


def chirp_generate(config: Optional[ChirpConfig] = None, device=None, ctx: Optional[Context] = None) -> torch.Tensor:
    """ChirpSync::generate (dual chirp) on the device, CUDA fp32 [n]."""
    config = config or ChirpConfig.default()
    n = lib().ria_chirp_generate_samples(C.addressof(config))
    if n <= 0:
        raise RiaError("bad chirp configuration")
    ctx = ctx or default_context()
    out = torch.empty(n, dtype=torch.float32, device=device or torch.device("cuda", ctx.device))
    ctx.set_stream(torch.cuda.current_stream(out.device))
    ctx.check(min(0, lib().ria_chirp_generate_dev(ctx.handle, C.addressof(config), _ptr(out), n)))
    return out

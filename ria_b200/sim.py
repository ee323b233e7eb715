"""Host-side mirror of the reference channel simulator (src/sim/hf_channel.hpp,
tools/cli_simulator.cpp:103-367) over the on-device channel kernels."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from ._lib import Context, RiaError, lib
from .fec import default_context


def awgn_batch(tx_pool: torch.Tensor, n_frames: int, snr_db, seed: int = 1, first_frame_id: int = 0,
               out: Optional[torch.Tensor] = None, ctx: Optional[Context] = None) -> torch.Tensor:
    """SimulatedChannel::applyChannel (AWGN) for a batch: frame f = tx_pool[(first+f) % P] + noise.

    tx_pool: CUDA fp32 [P, frame_len]; snr_db: float or CUDA fp32 [n_frames]."""
    if not (isinstance(tx_pool, torch.Tensor) and tx_pool.is_cuda and tx_pool.dtype == torch.float32):
        raise RiaError("awgn_batch wants a CUDA fp32 pool (no CPU fallback)")
    tx_pool = tx_pool.contiguous()
    P, L = tx_pool.shape
    if out is None:
        out = torch.empty((n_frames, L), dtype=torch.float32, device=tx_pool.device)
    ctx = ctx or default_context()
    ctx.set_stream(torch.cuda.current_stream(tx_pool.device))
    per = snr_db if isinstance(snr_db, torch.Tensor) else None
    ctx.check(lib().ria_channel_awgn_batch_dev(
        ctx.handle, C.c_void_p(tx_pool.data_ptr()), P, L,
        C.c_void_p(per.data_ptr()) if per is not None else C.c_void_p(0),
        float(0.0 if per is not None else snr_db), int(seed), int(first_frame_id), int(n_frames),
        C.c_void_p(out.data_ptr()), out.stride(0)))
    return out


class WattersonConfig(C.Structure):
    """ria_watterson_config (include/ria_b200.h) = sim::WattersonChannel::Config."""
    _fields_ = [("snr_db", C.c_float), ("delay_spread_ms", C.c_float), ("doppler_spread_hz", C.c_float),
                ("path1_gain", C.c_float), ("path2_gain", C.c_float), ("sample_rate", C.c_uint32),
                ("fading_enabled", C.c_uint32), ("multipath_enabled", C.c_uint32),
                ("noise_enabled", C.c_uint32), ("stationary_start", C.c_uint32)]

    AWGN, GOOD, MODERATE, POOR, FLUTTER = range(5)

    @classmethod
    def preset(cls, condition: int, snr_db: float = 20.0):
        """itu_r_f1487::{awgn, good, moderate, poor, flutter} (src/sim/hf_channel.hpp:411-488)."""
        c = cls()
        if lib().ria_watterson_preset(int(condition), float(snr_db), C.addressof(c)) != 0:
            raise ValueError("unknown channel condition")
        return c


def watterson_batch(cfg: WattersonConfig, tx_pool: torch.Tensor, n_frames: int, snr_db=None, seed: int = 1,
                    first_frame_id: int = 0, out: Optional[torch.Tensor] = None,
                    ctx: Optional[Context] = None) -> torch.Tensor:
    """WattersonChannel::process for a batch; snr_db: None (cfg.snr_db) or CUDA fp32 [n_frames]."""
    if not (isinstance(tx_pool, torch.Tensor) and tx_pool.is_cuda and tx_pool.dtype == torch.float32):
        raise RiaError("watterson_batch wants a CUDA fp32 pool (no CPU fallback)")
    tx_pool = tx_pool.contiguous()
    P, L = tx_pool.shape
    if out is None:
        out = torch.empty((n_frames, L), dtype=torch.float32, device=tx_pool.device)
    ctx = ctx or default_context()
    ctx.set_stream(torch.cuda.current_stream(tx_pool.device))
    ctx.check(lib().ria_channel_watterson_batch_dev(
        ctx.handle, C.addressof(cfg), C.c_void_p(tx_pool.data_ptr()), P, L,
        C.c_void_p(snr_db.data_ptr()) if snr_db is not None else C.c_void_p(0),
        int(seed), int(first_frame_id), int(n_frames), C.c_void_p(out.data_ptr()), out.stride(0)))
    return out

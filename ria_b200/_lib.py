"""Loader for libria_b200.so (the C ABI declared in include/ria_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C ria_b200/csrc``.  There is
no CPU fallback: if the shared object is missing or no CUDA device is usable, every compute
entry point raises ``RiaError``.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libria_b200.so")


DECODE_RETRY_LADDER = 1        # RIA_DECODE_* (include/ria_b200.h)
DECODE_FP_REPAIR = 2
DECODE_FULL = 3


class RiaError(RuntimeError):
    pass


_lib = None

_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> (restype, argtypes); mirrors include/ria_b200.h one to one.
_SIGNATURES = {
    "ria_version": (C.c_char_p, []),
    "ria_ctx_create": (_i32, [_i32, C.POINTER(_vp)]),
    "ria_ctx_destroy": (_i32, [_vp]),
    "ria_ctx_set_stream": (_i32, [_vp, _vp]),
    "ria_ctx_synchronize": (_i32, [_vp]),
    "ria_last_error": (C.c_char_p, [_vp]),
    "ria_ctx_launch_count": (_i64, [_vp]),
    "ria_ctx_set_timing": (_i32, [_vp, _i32]),
    "ria_ctx_set_decode_flags": (_i32, [_vp, _i32]),
    "ria_ctx_get_decode_flags": (_i32, [_vp]),
    "ria_ctx_get_timing": (_i32, [_vp, _i32, C.POINTER(C.c_double), C.POINTER(_i64)]),
    "ria_ldpc_params": (_i32, [_i32, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32)]),
    "ria_ldpc_get_matrix": (_i32, [_i32, _vp, _vp]),
    "ria_ldpc_decode_batch_dev": (_i32, [_vp, _i32, _i32, _f32, _vp, _i64, _vp, _i32, _vp, _vp]),
    "ria_ldpc_decode_batch_host": (_i32, [_vp, _i32, _i32, _f32, _vp, _i64, _vp, _i32, _vp, _vp]),
    "ria_ldpc_robust_decode_batch_dev": (_i32, [_vp, _i32, _vp, _i64, _vp, _i32, _vp, _vp, _vp]),
    "ria_ldpc_ladder_perturb_dev": (_i32, [_vp, _vp, _i64, _i32, _vp]),
    "ria_modem_config_for": (_i32, [_i32, _i32, _vp]),
    "ria_ofdm_symbol_samples": (_i32, [_vp]),
    "ria_ofdm_data_carriers": (_i32, [_vp]),
    "ria_ofdm_pilot_carriers": (_i32, [_vp]),
    "ria_ofdm_presynced_batch_dev": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _i64,
                                            _vp, _i32, _vp, _vp, _vp, _vp]),
    "ria_ofdm_presynced_batch_host": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _i64,
                                             _vp, _i32, _vp, _vp, _vp, _vp]),
    "ria_ofdm_presynced_batch_taps_dev": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _i64,
                                                 _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ria_frame_decode_batch_dev": (_i32, [_vp, _i32, _i32, _i32, _vp, _i32, _i64, _vp, _vp]),
    "ria_frame_decode_batch_host": (_i32, [_vp, _i32, _i32, _i32, _vp, _i32, _i64, _vp, _vp]),
    "ria_ofdm_rx_frames_dev": (_i32, [_vp, _vp, _i32, _i32, _vp, _i64, _i32, _vp, _vp, _i64,
                                      _vp, _vp, _vp]),
    "ria_ofdm_rx_frames_host": (_i32, [_vp, _vp, _i32, _i32, _vp, _i64, _i32, _vp, _vp, _i64,
                                       _vp, _vp, _vp]),
    "ria_encode_fixed_frame_batch_dev": (_i32, [_vp, _i32, _i32, _i32, _vp, _i64, _i32, _i64, _vp]),
    "ria_ofdm_tx_frame_samples": (_i32, [_vp, _i32]),
    "ria_ofdm_tx_frames_dev": (_i32, [_vp, _vp, _vp, _i64, _i32, _i64, _vp, _i64]),
    "ria_ofdm_cox_tx_frame_samples": (_i32, [_vp, _i32]),
    "ria_ofdm_cox_tx_frames_dev": (_i32, [_vp, _vp, _vp, _i64, _i32, _i64, _vp, _i64]),
    "ria_ofdm_cox_search_sync_batch_dev": (_i32, [_vp, _vp, _vp, _i64, _i32, _f32, _vp, _i64, _vp]),
    "ria_ofdm_cox_search_sync_batch_host": (_i32, [_vp, _vp, _vp, _i64, _i32, _f32, _vp, _i64, _vp]),
    "ria_ofdm_cox_correlation_batch_dev": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _i64, _vp]),
    "ria_ping_energy_batch_dev": (_i32, [_vp, _vp, _i64, _i32, _i32, _i64, _vp]),
    "ria_mcdpsk_zc_rx_frames_dev": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _f32, C.c_uint32, _i64, _i32, _i32, _f32,
                                        _vp, _i32, _vp, _i32, _vp, _vp, _vp]),
    "ria_ofdm_cox_rx_frames_dev": (_i32, [_vp, _vp, _i32, _i32, _vp, _i64, _i32, _i32, _f32, _vp, _i64, _vp, _vp, _vp, _vp]),
    "ria_ofdm_cox_rx_frames_host": (_i32, [_vp, _vp, _i32, _i32, _vp, _i64, _i32, _i32, _f32, _vp, _i64, _vp, _vp, _vp, _vp]),
    "ria_burst_deinterleave_batch_dev": (_i32, [_vp, _vp, _i32, _i32, _i64, _vp, _i32]),
    "ria_zc_config_default": (_i32, [_vp]),
    "ria_zc_detect_batch_dev": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _f32, C.c_uint32, _i64, _vp]),
    "ria_ofdm_data_sync_batch_dev": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _f32, _i64, _vp]),
    "ria_chirp_config_default": (_i32, [_vp]),
    "ria_chirp_detect_dual_batch_dev": (_i32, [_vp, _vp, _vp, _i64, _i32, _f32, _i64, _vp]),
    "ria_mcdpsk_soft_bits_per_frame": (_i32, [_vp, _i32]),
    "ria_mcdpsk_process_batch_dev": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _i64,
                                            _vp, _i32, _vp, _vp, _vp]),
    "ria_mcdpsk_process_batch_at_dev": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _i64,
                                               _vp, _i32, _vp, _vp, _vp]),
    "ria_mcdpsk_rx_frames_dev": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _f32, _i64, _i32, _i32, _f32,
                                        _vp, _i32, _vp, _i32, _vp, _vp, _vp]),
    "ria_mcdpsk_rx_frames_host": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _f32, _i64, _i32, _i32, _f32,
                                         _vp, _i32, _vp, _vp, _vp]),
    "ria_zc_preamble_host": (_i32, [_vp, _i32, _vp, _i32]),
    "ria_chirp_generate_host": (_i32, [_vp, _vp, _i32]),
    "ria_zc_preamble_samples": (_i32, [_vp]),
    "ria_chirp_generate_samples": (_i32, [_vp]),
    "ria_zc_preamble_dev": (_i32, [_vp, _vp, _i32, _vp, _i32]),
    "ria_chirp_generate_dev": (_i32, [_vp, _vp, _vp, _i32]),
    "ria_mcdpsk_tx_frame_samples": (_i32, [_vp, _i32]),
    "ria_mcdpsk_tx_frames_dev": (_i32, [_vp, _vp, _vp, _i64, _i32, _i64, _vp, _i64]),
    "ria_chase_combine_batch_dev": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i64]),
    "ria_recommend_waveform": (_i32, [_f32, _f32, _vp]),
    "ria_recommend_data_mode": (_i32, [_f32, _i32, _f32, _vp]),
    "ria_channel_awgn_batch_dev": (_i32, [_vp, _vp, _i32, _i32, _vp, _f32, C.c_uint64, _i64, _i64, _vp, _i64]),
    "ria_watterson_preset": (_i32, [_i32, _f32, _vp]),
    "ria_channel_watterson_batch_dev": (_i32, [_vp, _vp, _vp, _i32, _i32, _vp, C.c_uint64, _i64, _i64, _vp, _i64]),
    "ria_chirp_detect_dual_batch_host": (_i32, [_vp, _vp, _vp, _i64, _i32, _f32, _i64, _vp]),
    "ria_zc_detect_batch_host": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _f32, C.c_uint32, _i64, _vp]),
    "ria_ofdm_data_sync_batch_host": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _f32, _i64, _vp]),
    "ria_mcdpsk_process_batch_host": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _i64, _vp, _i32, _vp, _vp, _vp]),
    "ria_frame_counters_dev": (_i32, [_vp, _vp, _i64, _vp]),
    "ria_nccl_get_unique_id": (_i32, [_vp]),
    "ria_nccl_comm_create": (_i32, [_vp, _vp, _i32, _i32, _vp]),
    "ria_nccl_comm_destroy": (_i32, [_vp]),
    "ria_counters_allreduce": (_i32, [_vp, _vp, _vp, _i32]),
    "ria_crc16": (C.c_uint16, [_vp, C.c_size_t]),
    "ria_channel_interleaver_step": (_i32, [_i32, _i32]),
}


def exported_symbols():
    return sorted(_SIGNATURES)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RiaError(
                f"{LIB_PATH} is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C ria_b200/csrc`). ria_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)      # AttributeError here = header/library mismatch: fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class Context:
    """One ``ria_ctx`` per GPU (include/ria_b200.h)."""

    def __init__(self, device: int = 0, stream=None):
        self._L = lib()
        h = _vp()
        rc = self._L.ria_ctx_create(int(device), C.byref(h))
        if rc != 0 or not h.value:
            raise RiaError(f"ria_ctx_create(device={device}) failed with {rc}: no usable B200; "
                           "ria_b200 has no CPU fallback")
        self.handle = h
        self.device = device
        if stream is not None:
            self.set_stream(stream)

    def set_stream(self, stream) -> None:
        """Bind to a torch.cuda.Stream (or raw cudaStream_t integer)."""
        ptr = getattr(stream, "cuda_stream", stream)
        self.check(self._L.ria_ctx_set_stream(self.handle, _vp(ptr)))

    def synchronize(self) -> None:
        self.check(self._L.ria_ctx_synchronize(self.handle))

    @property
    def launch_count(self) -> int:
        return int(self._L.ria_ctx_launch_count(self.handle))

    def set_decode_flags(self, flags: int) -> None:
        """RIA_DECODE_RETRY_LADDER (1): frame entry points run v2::decodeFixedFrame's retry ladder."""
        self.check(self._L.ria_ctx_set_decode_flags(self.handle, int(flags)))

    def get_decode_flags(self) -> int:
        return int(self._L.ria_ctx_get_decode_flags(self.handle))

    def set_timing(self, enable: bool) -> None:
        self.check(self._L.ria_ctx_set_timing(self.handle, int(enable)))

    def get_timing(self, kind: int):
        """(total_ms, launches) of one kernel kind since set_timing(True)."""
        ms, n = C.c_double(0), _i64(0)
        self.check(self._L.ria_ctx_get_timing(self.handle, int(kind), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def check(self, rc: int) -> None:
        if rc != 0:
            msg = self._L.ria_last_error(self.handle)
            raise RiaError(f"ria_b200 error {rc}: {msg.decode() if msg else ''}")

    def close(self) -> None:
        if getattr(self, "handle", None) is not None and self.handle.value:
            self._L.ria_ctx_destroy(self.handle)
            self.handle = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

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


class ModemConfig(C.Structure):
    """Same layout as ria_modem_config (include/ria_b200.h)."""
    _fields_ = [(n, C.c_uint32) for n in (
        "sample_rate", "center_freq", "fft_size", "num_carriers", "cp_mode", "symbol_guard",
        "use_pilots", "pilot_spacing", "modulation", "training_symbols")]

    @classmethod
    def make(cls, modulation, pilot_spacing, use_pilots=1, cp_mode=1, num_carriers=59):
        return cls(48000, 1500, 1024, num_carriers, cp_mode, 0, use_pilots, pilot_spacing,
                   modulation, 2)

    def symbol_samples(self):
        return self.fft_size + (32, 48, 64)[self.cp_mode] * (self.fft_size // 512) + self.symbol_guard

    def pilots(self):
        if not self.use_pilots:
            return 0
        return (self.num_carriers + self.pilot_spacing - 1) // self.pilot_spacing

    def data_carriers(self):
        return self.num_carriers - self.pilots()


class StreamDecodeResult(C.Structure):
    """gui::DecodeResult fields (src/gui/modem/streaming_decoder.hpp:70-79)"""
    _fields_ = [("success", C.c_int32), ("frame_type", C.c_int32), ("codewords_ok", C.c_int32),
                ("codewords_failed", C.c_int32), ("is_ping", C.c_int32), ("n_bytes", C.c_int32)]


class StreamStepResult(C.Structure):
    """result of one StreamingDecoder::decodeCurrentFrame step (oracle/ref_shim.cpp ref_stream_step)"""
    _fields_ = [("state", C.c_int32), ("pending_total_cw", C.c_int32), ("has_frame", C.c_int32),
                ("frame", StreamDecodeResult), ("last_cfo", C.c_float), ("sync_pos", C.c_int32)]


class McdpskConfig(C.Structure):
    """Same layout as ria_mcdpsk_config (include/ria_b200.h)."""
    _fields_ = [("sample_rate", C.c_float), ("num_carriers", C.c_uint32), ("freq_low", C.c_float),
                ("freq_high", C.c_float), ("samples_per_symbol", C.c_uint32),
                ("bits_per_symbol", C.c_uint32), ("spreading", C.c_uint32),
                ("training_symbols", C.c_uint32)]

    @classmethod
    def make(cls, bits=1, spreading=4, carriers=10, training=8):
        return cls(48000.0, carriers, 500.0, 2500.0, 512, bits, spreading, training)


class ZcConfig(C.Structure):
    """Same layout as ria_zc_config (include/ria_b200.h)."""
    _fields_ = [("sample_rate", C.c_float), ("sequence_length", C.c_int32), ("upsample_factor", C.c_int32),
                ("num_repetitions", C.c_int32), ("carrier_freq", C.c_float), ("gap_ms", C.c_float),
                ("root_ping", C.c_int32), ("root_pong", C.c_int32), ("root_data", C.c_int32),
                ("root_control", C.c_int32)]

    @classmethod
    def default(cls):
        return cls(48000.0, 127, 8, 2, 1500.0, 10.0, 1, 3, 5, 7)


class SyncResult(C.Structure):
    """Same layout as ria_sync_result (include/ria_b200.h)."""
    _fields_ = [("detected", C.c_int32), ("start_sample", C.c_int32), ("correlation", C.c_float),
                ("cfo_hz", C.c_float), ("snr_estimate", C.c_float), ("root", C.c_int32),
                ("frame_type", C.c_int32), ("aux", C.c_int32)]


class WattersonConfig(C.Structure):
    """Same layout as ria_watterson_config."""
    _fields_ = [("snr_db", C.c_float), ("delay_spread_ms", C.c_float), ("doppler_spread_hz", C.c_float),
                ("path1_gain", C.c_float), ("path2_gain", C.c_float), ("sample_rate", C.c_uint32),
                ("fading_enabled", C.c_uint32), ("multipath_enabled", C.c_uint32),
                ("noise_enabled", C.c_uint32), ("stationary_start", C.c_uint32)]


class WaveformRecommendation(C.Structure):
    _fields_ = [("waveform", C.c_int32), ("modulation", C.c_int32), ("rate", C.c_int32),
                ("estimated_throughput_bps", C.c_float), ("num_carriers", C.c_int32), ("spreading", C.c_int32)]


class FrameStatus(C.Structure):
    """Same layout as ria_frame_status (include/ria_b200.h)."""
    _fields_ = [("cw_ok", C.c_uint8 * 4), ("cw_iters", C.c_int32 * 4), ("all_ok", C.c_uint8),
                ("header_valid", C.c_uint8), ("frame_crc_ok", C.c_uint8), ("type", C.c_uint8),
                ("seq", C.c_uint16), ("payload_len", C.c_uint16), ("src_hash", C.c_uint32),
                ("dst_hash", C.c_uint32), ("total_cw", C.c_uint8), ("ladder_cw_mask", C.c_uint8),
                ("ladder_max_attempt", C.c_uint8), ("fp_repair", C.c_uint8)]


# ultra::Modulation (include/ultra/types.hpp:27-39)
DBPSK, BPSK, DQPSK, QPSK, D8PSK, QAM8, QAM16, QAM32, QAM64 = range(9)
QAM256 = 10
BITS_PER_CARRIER = {DBPSK: 1, BPSK: 1, DQPSK: 2, QPSK: 2, D8PSK: 3, QAM8: 3, QAM16: 4, QAM32: 5, QAM64: 6, QAM256: 8}
BYTES_PER_CW = {R1_4: 20, R1_2: 40, R2_3: 54, R3_4: 60, R5_6: 67}


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
        L.orc_cox_correlation.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int] + [C.POINTER(C.c_float)] * 5
        L.orc_cox_correlation.restype = C.c_int
        L.orc_cox_coarse_cfo.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint]
        L.orc_cox_coarse_cfo.restype = C.c_float
        self._codes = {}

    def code(self, rate: int) -> _Code:
        if rate not in self._codes:
            c = _Code()
            self.lib.orc_ldpc_build(rate, C.byref(c))
            self._codes[rate] = c
        return self._codes[rate]

    def cox_correlation(self, samples, offset: int, cp_len: int, fft_len: int = 1024):
        """Schmidl-Cox metric at `offset` (ofdm_sync.cpp:118-163) -> (metric, P, R1, R2)"""
        samples = np.ascontiguousarray(samples, dtype=np.float32)
        m, pr, pi, r1, r2 = (C.c_float() for _ in range(5))
        rc = self.lib.orc_cox_correlation(samples, len(samples), int(offset), int(cp_len), int(fft_len),
                                          C.byref(m), C.byref(pr), C.byref(pi), C.byref(r1), C.byref(r2))
        assert rc >= 0
        return np.float32(m.value), complex(pr.value, pi.value), np.float32(r1.value), np.float32(r2.value)

    def cox_coarse_cfo(self, samples, sync_offset: int, cp_len: int, fft_len: int = 1024, sample_rate: int = 48000):
        """Impl::estimateCoarseCFO (ofdm_sync.cpp:230-261)"""
        samples = np.ascontiguousarray(samples, dtype=np.float32)
        return np.float32(self.lib.orc_cox_coarse_cfo(samples, len(samples), int(sync_offset), int(cp_len), int(fft_len), int(sample_rate)))

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
        cfgp = C.POINTER(ModemConfig)
        L.ref_ofdm_tx_frame.argtypes = [cfgp, _u8p, C.c_int, _f32p, C.c_int]
        L.ref_ofdm_tx_frame.restype = C.c_int
        L.ref_ofdm_demod_new.argtypes = [cfgp]
        L.ref_ofdm_demod_new.restype = C.c_void_p
        L.ref_ofdm_demod_free.argtypes = [C.c_void_p]
        fp = C.POINTER(C.c_float)
        L.ref_ofdm_process_presynced.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_float, C.c_float,
                                                 _f32p, C.c_int, C.POINTER(C.c_int), fp, fp, fp, C.c_void_p]
        L.ref_ofdm_process_presynced.restype = C.c_int
        L.ref_fft_forward.argtypes = [C.c_int, _f32p, _f32p]
        L.ref_ofdm_symbol_bins.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_float, C.c_float, _f32p]
        L.ref_encode_fixed_frame.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, _u8p, C.c_int]
        L.ref_encode_fixed_frame.restype = C.c_int
        L.ref_frame_decode_first_pass.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, _u8p, _u8p, _i32p]
        L.ref_decode_fixed_frame_full.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, _u8p, _u8p]
        L.ref_burst_deinterleave.argtypes = [_f32p, C.c_int, _f32p]
        L.ref_ladder_perturb.argtypes = [_f32p, C.c_int, C.c_uint, C.c_float, C.c_int, _f32p]
        L.ref_parse_header.argtypes = [_u8p, C.c_int, C.POINTER(FrameStatus)]
        L.ref_frame_status_reassembled.argtypes = [_u8p, _u8p, C.c_int, C.POINTER(FrameStatus)]
        L.ref_stream_decoder_new.restype = C.c_void_p
        L.ref_stream_decoder_free.argtypes = [C.c_void_p]
        L.ref_stream_decode_mcdpsk_frame.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, C.POINTER(StreamDecodeResult), _u8p, C.c_int]
        L.ref_encode_frame_with_ldpc.argtypes = [_u8p, C.c_int, C.c_int, _u8p, C.c_int]
        L.ref_stream_decode_ofdm_frame.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                   C.POINTER(StreamDecodeResult), _u8p, C.c_int]
        L.ref_encode_frame_with_ldpc.restype = C.c_int
        L.ref_crc16.argtypes = [_u8p, C.c_int]
        L.ref_crc16.restype = C.c_uint16
        L.ref_make_data_frame.argtypes = [C.c_char_p, C.c_char_p, C.c_int, _u8p, C.c_int, _u8p, C.c_int]
        L.ref_make_data_frame.restype = C.c_int
        self._demods = {}
        L.ref_watterson_process.argtypes = [C.POINTER(WattersonConfig), C.c_uint, _f32p, C.c_int, _f32p]
        L.ref_recommend_waveform.argtypes = [C.c_float, C.c_float, C.POINTER(WaveformRecommendation)]
        L.ref_recommend_data_mode.argtypes = [C.c_float, C.c_int, C.c_float, C.POINTER(WaveformRecommendation)]
        L.ref_chase_combine.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, _f32p, C.POINTER(C.c_int)]
        L.ref_chase_combine.restype = C.c_int
        L.ref_ofdm_data_sync.argtypes = [C.POINTER(ModemConfig), _f32p, C.c_int, C.c_float, C.c_float, C.POINTER(SyncResult)]
        L.ref_stream_setup_ofdm.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.ref_stream_setup_mcdpsk.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ref_stream_min_control_samples.argtypes = [C.c_void_p]
        L.ref_stream_min_control_samples.restype = C.c_int
        L.ref_stream_step.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_float,
                                      C.POINTER(StreamStepResult), _u8p, C.c_int]
        L.ref_stream_encode.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _u8p, C.c_int, _f32p, C.c_int]
        L.ref_stream_encode.restype = C.c_int
        L.ref_stream_encode_burst.argtypes = [C.c_int, C.c_int, C.c_int, _u8p, C.c_int, C.c_int, _f32p, C.c_int]
        L.ref_stream_encode_burst.restype = C.c_int
        L.ref_stream_burst_group.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int,
                                             C.POINTER(StreamDecodeResult), _u8p, C.c_int, C.c_int, C.POINTER(C.c_float)]
        L.ref_stream_burst_group.restype = C.c_int
        L.ref_make_ack_frame.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, _u8p, C.c_int]
        L.ref_make_ack_frame.restype = C.c_int
        L.ref_ofdm_cox_tx_frame.argtypes = [C.POINTER(ModemConfig), _u8p, C.c_int, _f32p, C.c_int]
        L.ref_ofdm_cox_tx_frame.restype = C.c_int
        L.ref_ofdm_cox_search_sync.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_float, C.POINTER(C.c_float),
                                               C.POINTER(C.c_longlong), C.POINTER(C.c_float)]
        L.ref_ofdm_cox_search_sync.restype = C.c_int
        L.ref_ofdm_cox_correlation.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int]
        L.ref_ofdm_cox_correlation.restype = C.c_float
        L.ref_ofdm_cox_refine_lts.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, C.POINTER(C.c_float)]
        L.ref_ofdm_cox_refine_lts.restype = C.c_longlong
        zcp = C.POINTER(ZcConfig)
        L.ref_zc_preamble.argtypes = [zcp, C.c_int, _f32p, C.c_int]
        L.ref_zc_preamble.restype = C.c_int
        L.ref_zc_detect.argtypes = [zcp, _f32p, C.c_int, C.c_float, C.c_uint, C.c_float, C.POINTER(SyncResult)]
        L.ref_chirp_generate.argtypes = [_f32p, C.c_int]
        L.ref_chirp_generate.restype = C.c_int
        L.ref_chirp_detect_dual.argtypes = [_f32p, C.c_int, C.c_float, C.POINTER(SyncResult)]
        mcp = C.POINTER(McdpskConfig)
        L.ref_mcdpsk_tx_frame.argtypes = [mcp, _u8p, C.c_int, _f32p, C.c_int]
        L.ref_mcdpsk_tx_frame.restype = C.c_int
        L.ref_mcdpsk_demod_new.argtypes = [mcp]
        L.ref_mcdpsk_demod_new.restype = C.c_void_p
        L.ref_mcdpsk_demod_free.argtypes = [C.c_void_p]
        L.ref_mcdpsk_process.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_float, C.c_float, _f32p, C.c_int,
                                         C.POINTER(C.c_int), fp, fp]
        L.ref_mcdpsk_process.restype = C.c_int

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


    # ---- OFDM ----
    def ofdm_tx_frame(self, cfg: ModemConfig, data) -> np.ndarray:
        data = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8))
        cap = cfg.symbol_samples() * (2 + len(data) * 8 // max(1, cfg.data_carriers()) + 4)
        out = np.zeros(cap, np.float32)
        n = self.lib.ref_ofdm_tx_frame(C.byref(cfg), data, len(data), out, cap)
        assert n >= 0, n
        return out[:n].copy()

    def ofdm_cox_tx_frame(self, cfg: ModemConfig, data) -> np.ndarray:
        """[guard][4 x STS][2 x LTS][data symbols] as OFDMNvisWaveform transmits it"""
        data = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8))
        cap = cfg.symbol_samples() * (8 + len(data) * 8 // max(1, cfg.data_carriers()) + 4)
        out = np.zeros(cap, np.float32)
        n = self.lib.ref_ofdm_cox_tx_frame(C.byref(cfg), data, len(data), out, cap)
        assert n >= 0, n
        return out[:n].copy()

    def ofdm_cox_search_sync(self, cfg: ModemConfig, samples, threshold=0.8, noise_floor=0.0):
        """OFDMDemodulator::searchForSync -> (found, lts_position or -1, cfo_hz, noise_floor_after)"""
        samples = np.ascontiguousarray(samples, dtype=np.float32)
        nf = C.c_float(noise_floor); pos = C.c_longlong(-1); cfo = C.c_float(0)
        f = self.lib.ref_ofdm_cox_search_sync(self._demod(cfg), samples, len(samples), threshold,
                                              C.byref(nf), C.byref(pos), C.byref(cfo))
        return bool(f), pos.value, cfo.value, nf.value

    def ofdm_cox_correlation(self, cfg: ModemConfig, samples, offset) -> float:
        samples = np.ascontiguousarray(samples, dtype=np.float32)
        return self.lib.ref_ofdm_cox_correlation(self._demod(cfg), samples, len(samples), int(offset))

    def ofdm_cox_refine_lts(self, cfg: ModemConfig, samples, coarse_sts):
        samples = np.ascontiguousarray(samples, dtype=np.float32)
        cfo = C.c_float(0)
        r = self.lib.ref_ofdm_cox_refine_lts(self._demod(cfg), samples, len(samples), int(coarse_sts), C.byref(cfo))
        return r, cfo.value

    def _demod(self, cfg: ModemConfig):
        key = bytes(cfg)
        if key not in self._demods:
            self._demods[key] = self.lib.ref_ofdm_demod_new(C.byref(cfg))
        return self._demods[key]

    def ofdm_process_presynced(self, cfg: ModemConfig, samples, cfo_hz=0.0, phase=0.0):
        """-> dict(ready, soft, snr_db, cfo, fading, h)   (one frame)"""
        samples = np.ascontiguousarray(samples, dtype=np.float32)
        cap = 8192
        soft = np.zeros(cap, np.float32)
        n_soft = C.c_int(0)
        snr, cfo, fad = C.c_float(0), C.c_float(0), C.c_float(0)
        h = np.zeros((cfg.num_carriers, 2), np.float32)
        ready = self.lib.ref_ofdm_process_presynced(
            self._demod(cfg), samples, len(samples), cfo_hz, phase, soft, cap, C.byref(n_soft),
            C.byref(snr), C.byref(cfo), C.byref(fad), h.ctypes.data)
        return dict(ready=bool(ready), soft=soft[: n_soft.value].copy(), snr_db=snr.value,
                    cfo=cfo.value, fading=fad.value, h=h[:, 0] + 1j * h[:, 1])

    def fft_forward(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.complex64)
        out = np.zeros_like(x)
        self.lib.ref_fft_forward(len(x), x.view(np.float32), out.view(np.float32))
        return out

    def ofdm_symbol_bins(self, cfg: ModemConfig, samples, n_sym, cfo_hz=0.0, phase=0.0):
        samples = np.ascontiguousarray(samples, dtype=np.float32)
        bins = np.zeros((n_sym, cfg.num_carriers, 2), np.float32)
        self.lib.ref_ofdm_symbol_bins(self._demod(cfg), samples, n_sym, cfo_hz, phase, bins)
        return bins[..., 0] + 1j * bins[..., 1]

    # ---- channel / chase / selection ----
    def watterson_process(self, cfg: WattersonConfig, seed: int, x) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.zeros_like(x)
        self.lib.ref_watterson_process(C.byref(cfg), seed, x, len(x), out)
        return out

    def recommend_waveform(self, snr_db, fading) -> WaveformRecommendation:
        r = WaveformRecommendation()
        self.lib.ref_recommend_waveform(snr_db, fading, C.byref(r))
        return r

    def recommend_data_mode(self, snr_db, waveform, fading) -> WaveformRecommendation:
        r = WaveformRecommendation()
        self.lib.ref_recommend_data_mode(snr_db, waveform, fading, C.byref(r))
        return r

    def chase_combine(self, soft, cw_index=1, total_cw=4):
        soft = np.ascontiguousarray(soft, dtype=np.float32).reshape(-1, 648)
        out = np.zeros(648, np.float32)
        cnt = C.c_int(0)
        stored = self.lib.ref_chase_combine(soft, soft.shape[0], cw_index, total_cw, out, C.byref(cnt))
        return out, stored, cnt.value

    # ---- sync ----
    def ofdm_data_sync(self, cfg: ModemConfig, samples, known_cfo=0.0, threshold=0.3) -> SyncResult:
        samples = np.ascontiguousarray(samples, dtype=np.float32)
        r = SyncResult()
        self.lib.ref_ofdm_data_sync(C.byref(cfg), samples, len(samples), known_cfo, threshold, C.byref(r))
        return r

    def zc_preamble(self, cfg: ZcConfig, frame_type: int) -> np.ndarray:
        out = np.zeros(8192, np.float32)
        n = self.lib.ref_zc_preamble(C.byref(cfg), frame_type, out, len(out))
        assert n > 0
        return out[:n].copy()

    def zc_detect(self, cfg: ZcConfig, samples, threshold=0.3, root_mask=0xF, known_cfo=0.0) -> SyncResult:
        samples = np.ascontiguousarray(samples, dtype=np.float32)
        r = SyncResult()
        self.lib.ref_zc_detect(C.byref(cfg), samples, len(samples), threshold, root_mask, known_cfo, C.byref(r))
        return r

    def chirp_generate(self) -> np.ndarray:
        out = np.zeros(60000, np.float32)
        n = self.lib.ref_chirp_generate(out, len(out))
        assert n > 0
        return out[:n].copy()

    def chirp_detect_dual(self, samples, threshold=0.15) -> SyncResult:
        samples = np.ascontiguousarray(samples, dtype=np.float32)
        r = SyncResult()
        self.lib.ref_chirp_detect_dual(samples, len(samples), threshold, C.byref(r))
        return r

    # ---- MC-DPSK ----
    def mcdpsk_tx_frame(self, cfg: McdpskConfig, data) -> np.ndarray:
        data = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8))
        bits_sym = cfg.num_carriers * cfg.bits_per_symbol
        cap = 512 * (cfg.training_symbols + 2 + (len(data) * 8 // bits_sym + 2) * cfg.spreading)
        out = np.zeros(cap, np.float32)
        n = self.lib.ref_mcdpsk_tx_frame(C.byref(cfg), data, len(data), out, cap)
        assert n >= 0, n
        return out[:n].copy()

    def mcdpsk_process(self, cfg: McdpskConfig, samples, cfo_hz=0.0, phase=0.0):
        key = ("mc", bytes(cfg))
        if key not in self._demods:
            self._demods[key] = self.lib.ref_mcdpsk_demod_new(C.byref(cfg))
        samples = np.ascontiguousarray(samples, dtype=np.float32)
        cap = 16384
        soft = np.zeros(cap, np.float32)
        n_soft = C.c_int(0)
        fad, cfo = C.c_float(0), C.c_float(0)
        ready = self.lib.ref_mcdpsk_process(self._demods[key], samples, len(samples), cfo_hz, phase,
                                            soft, cap, C.byref(n_soft), C.byref(fad), C.byref(cfo))
        return dict(ready=bool(ready), soft=soft[: n_soft.value].copy(), fading=fad.value, cfo=cfo.value)

    # ---- fixed frame ----
    def encode_fixed_frame(self, data, rate, use_ci, bps) -> np.ndarray:
        data = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8))
        out = np.zeros(400, np.uint8)
        n = self.lib.ref_encode_fixed_frame(data, len(data), rate, int(use_ci), bps, out, len(out))
        assert n == 324, n
        return out[:n].copy()

    def frame_decode_first_pass(self, soft, rate, use_ci, bps):
        soft = np.ascontiguousarray(soft, dtype=np.float32)
        assert soft.size >= 2592
        data = np.zeros(4 * BYTES_PER_CW[rate], np.uint8)
        ok = np.zeros(4, np.uint8)
        iters = np.zeros(4, np.int32)
        self.lib.ref_frame_decode_first_pass(soft, rate, int(use_ci), bps, data, ok, iters)
        return data, ok, iters

    def decode_fixed_frame_full(self, soft, rate, use_ci, bps):
        soft = np.ascontiguousarray(soft, dtype=np.float32)
        data = np.zeros(4 * BYTES_PER_CW[rate], np.uint8)
        ok = np.zeros(4, np.uint8)
        self.lib.ref_decode_fixed_frame_full(soft, soft.size, rate, int(use_ci), bps, data, ok)
        return data, ok

    def burst_deinterleave(self, physical) -> np.ndarray:
        """fec::BurstInterleaver::deinterleave on [n, 2592] soft bits."""
        physical = np.ascontiguousarray(physical, dtype=np.float32)
        out = np.empty_like(physical)
        self.lib.ref_burst_deinterleave(physical, physical.shape[0], out)
        return out

    def ladder_perturb(self, llr, seed: int, sigma: float, kind: int) -> np.ndarray:
        """std::mt19937(seed) + std::normal_distribution<float>(0, sigma) perturbation (frame_v2.cpp:1426-1542)."""
        llr = np.ascontiguousarray(llr, dtype=np.float32)
        out = np.empty_like(llr)
        self.lib.ref_ladder_perturb(llr, llr.size, seed & 0xFFFFFFFF, sigma, kind, out)
        return out

    def parse_header(self, data) -> FrameStatus:
        data = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8))
        st = FrameStatus()
        self.lib.ref_parse_header(data, len(data), C.byref(st))
        return st

    def frame_status_reassembled(self, data, ok, bpc: int) -> FrameStatus:
        """header of codeword 0 + whether CodewordStatus::reassemble() gives a frame DataFrame::deserialize accepts"""
        data = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8))
        ok = np.ascontiguousarray(np.asarray(ok, dtype=np.uint8))
        assert len(data) >= 4 * bpc and len(ok) == 4
        st = FrameStatus()
        self.lib.ref_frame_status_reassembled(data, ok, int(bpc), C.byref(st))
        return st

    # ---- StreamingDecoder frame-level decode ----
    def stream_decoder(self):
        """a reference StreamingDecoder object (owns a HARQ chase cache); free with stream_decoder_free"""
        return self.lib.ref_stream_decoder_new()

    def stream_decoder_free(self, h):
        self.lib.ref_stream_decoder_free(h)

    def stream_decode_mcdpsk_frame(self, h, soft, rate: int):
        """StreamingDecoder::decodeMCDPSKFrame -> (StreamDecodeResult, frame bytes)"""
        soft = np.ascontiguousarray(soft, dtype=np.float32)
        res = StreamDecodeResult()
        buf = np.zeros(1024, np.uint8)
        self.lib.ref_stream_decode_mcdpsk_frame(h, soft, len(soft), int(rate), C.byref(res), buf, len(buf))
        return res, bytes(buf[: res.n_bytes])

    def stream_decode_ofdm_frame(self, h, soft, connected: bool, modulation: int, rate: int, data_carriers: int,
                                 use_channel_interleave: bool = True):
        """StreamingDecoder::decodeFrame for an OFDM receiver in that state -> (StreamDecodeResult, frame bytes)"""
        soft = np.ascontiguousarray(soft, dtype=np.float32)
        res = StreamDecodeResult()
        buf = np.zeros(1024, np.uint8)
        self.lib.ref_stream_decode_ofdm_frame(h, soft, len(soft), int(bool(connected)), int(modulation), int(rate),
                                              int(data_carriers), int(bool(use_channel_interleave)), C.byref(res), buf, len(buf))
        return res, bytes(buf[: res.n_bytes])

    def stream_setup_ofdm(self, h, connected: bool, modulation: int, rate: int):
        self.lib.ref_stream_setup_ofdm(C.c_void_p(h), int(bool(connected)), int(modulation), int(rate))

    def stream_setup_mcdpsk(self, h, connected: bool, carriers: int, modulation: int, rate: int, spreading: int):
        """spreading: SpreadingMode enum (0 NONE, 1 TIME_2X, 2 TIME_4X)"""
        self.lib.ref_stream_setup_mcdpsk(C.c_void_p(h), int(bool(connected)), int(carriers), int(modulation), int(rate), int(spreading))

    def stream_min_control_samples(self, h) -> int:
        return int(self.lib.ref_stream_min_control_samples(C.c_void_p(h)))

    def stream_step(self, h, samples, sync_pos=0, sync_cfo=0.0, sync_snr=10.0, pending_total_cw=0, last_cfo=0.0):
        """StreamingDecoder::decodeCurrentFrame with the ring buffer holding `samples` and sync found at sync_pos
        -> (StreamStepResult, frame bytes)"""
        samples = np.ascontiguousarray(samples, dtype=np.float32)
        res = StreamStepResult()
        buf = np.zeros(1024, np.uint8)
        self.lib.ref_stream_step(C.c_void_p(h), samples, len(samples), int(sync_pos), float(sync_cfo), float(sync_snr), int(pending_total_cw),
                                 float(last_cfo), C.byref(res), buf, len(buf))
        return res, bytes(buf[: res.frame.n_bytes])

    def stream_encode(self, waveform: int, modulation: int, rate: int, kind: int, frame=b"", carriers=10, spreading=0) -> np.ndarray:
        """StreamingEncoder::encodeFrame (kind 0) / encodeFrameLight (1) / encodePing (2); waveform 1 OFDM_CHIRP, 2 MC_DPSK"""
        frame = np.ascontiguousarray(np.frombuffer(bytes(frame), dtype=np.uint8)) if len(frame) else np.zeros(1, np.uint8)
        cap = 400000
        out = np.zeros(cap, np.float32)
        n = self.lib.ref_stream_encode(int(waveform), int(carriers), int(spreading), int(modulation), int(rate), int(kind),
                                       frame, len(frame), out, cap)
        assert n > 0, n
        return out[:n].copy()

    def stream_encode_burst(self, modulation: int, rate: int, frames, group: int = 4) -> np.ndarray:
        """StreamingEncoder::encodeBurstLight with burst interleaving on; frames: equally long byte strings"""
        flen = len(frames[0])
        buf = np.ascontiguousarray(np.frombuffer(b"".join(bytes(f) for f in frames), dtype=np.uint8))
        cap = 1200000
        out = np.zeros(cap, np.float32)
        n = self.lib.ref_stream_encode_burst(int(modulation), int(rate), int(group), buf, flen, len(frames), out, cap)
        assert n > 0, n
        return out[:n].copy()

    def stream_burst_group(self, h, samples, sync_pos, sync_cfo=0.0, last_cfo=0.0, group=4):
        """decodeCurrentFrame with the burst marker latched + accumulateBurstFrames to the end of the group
        -> (list of (StreamDecodeResult, frame bytes), last_cfo) or (None, last_cfo) when the group was not finished"""
        samples = np.ascontiguousarray(samples, dtype=np.float32)
        res = (StreamDecodeResult * 16)()
        buf = np.zeros((16, 1024), np.uint8)
        lc = C.c_float(0)
        n = self.lib.ref_stream_burst_group(C.c_void_p(h), samples, len(samples), int(sync_pos), float(sync_cfo), float(last_cfo),
                                            int(group), res, buf, 1024, 16, C.byref(lc))
        if n < 0:
            return None, lc.value
        return [(res[i], bytes(buf[i, : res[i].n_bytes])) for i in range(n)], lc.value

    def make_ack_frame(self, src: str, dst: str, seq: int, nack=False) -> bytes:
        out = np.zeros(64, np.uint8)
        n = self.lib.ref_make_ack_frame(src.encode(), dst.encode(), int(seq), int(bool(nack)), out, len(out))
        assert n > 0, n
        return bytes(out[:n])

    def encode_frame_with_ldpc(self, frame, rate: int) -> np.ndarray:
        """v2::encodeFrameWithLDPC -> coded bytes [n_cw, 81]"""
        frame = np.ascontiguousarray(np.frombuffer(bytes(frame), dtype=np.uint8))
        out = np.zeros((32, 81), np.uint8)
        n = self.lib.ref_encode_frame_with_ldpc(frame, len(frame), int(rate), out, 32)
        assert n > 0, n
        return out[:n].copy()

    def crc16(self, data) -> int:
        data = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8))
        return int(self.lib.ref_crc16(data, len(data)))

    def make_data_frame(self, src: str, dst: str, seq: int, payload) -> bytes:
        payload = np.ascontiguousarray(np.frombuffer(bytes(payload), dtype=np.uint8))
        out = np.zeros(len(payload) + 64, np.uint8)
        n = self.lib.ref_make_data_frame(src.encode(), dst.encode(), seq, payload, len(payload), out, len(out))
        assert n > 0
        return out[:n].tobytes()


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

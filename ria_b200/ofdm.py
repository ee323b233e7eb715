"""Host-side mirror of the reference OFDM receive interfaces over the C ABI.

  * ``ModemConfig``        <- ultra::ModemConfig, RX-relevant fields (include/ultra/types.hpp:193-289)
  * ``OFDMDemodulator``    <- ultra::OFDMDemodulator presynced path (include/ultra/ofdm.hpp:58-139,
                              src/ofdm/demodulator.cpp:1212-1221, 1250-1414)
  * ``OFDMChirpWaveform``  <- the RX half of ultra::OFDMChirpWaveform / IWaveform
                              (src/waveform/ofdm_chirp_waveform.cpp:79-105, 391-485,
                               src/waveform/waveform_interface.hpp:47-220)
  * ``decode_fixed_frame_batch`` <- first pass of v2::decodeFixedFrame
                              (src/protocol/frame_v2.cpp:1335-1385) + v2::parseHeader
  * ``OfdmRxChain``        <- the flat batched driver that replaces StreamingDecoder's per-frame
                              decodeCurrentFrame for a batch of presynced frames.

Batch methods take CUDA tensors (torch is only used for device memory and streams).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from ._lib import Context, RiaError, lib, DECODE_RETRY_LADDER, DECODE_FP_REPAIR
from .fec import code_params, default_context

# ultra::Modulation (include/ultra/types.hpp:27-39)
DBPSK, BPSK, DQPSK, QPSK, D8PSK, QAM8, QAM16, QAM32, QAM64 = range(9)
QAM256 = 10
_BITS = {DBPSK: 1, BPSK: 1, DQPSK: 2, QPSK: 2, D8PSK: 3, QAM8: 3, QAM16: 4, QAM32: 5, QAM64: 6, QAM256: 8}


def getBitsPerSymbol(mod: int) -> int:
    return _BITS.get(int(mod), 1)


class ModemConfig(C.Structure):
    """ria_modem_config (include/ria_b200.h)."""
    _fields_ = [(n, C.c_uint32) for n in (
        "sample_rate", "center_freq", "fft_size", "num_carriers", "cp_mode", "symbol_guard",
        "use_pilots", "pilot_spacing", "modulation", "training_symbols")]

    @classmethod
    def default(cls, modulation=QPSK, use_pilots=0, pilot_spacing=2, cp_mode=1, num_carriers=59):
        """ModemConfig{} defaults of the reference (1024 FFT, 59 carriers, MEDIUM CP)."""
        return cls(48000, 1500, 1024, num_carriers, cp_mode, 0, use_pilots, pilot_spacing, modulation, 2)

    @classmethod
    def for_waveform(cls, modulation: int, rate: int):
        """What OFDMChirpWaveform::configure(mod, rate) produces."""
        c = cls()
        if lib().ria_modem_config_for(int(modulation), int(rate), C.addressof(c)) != 0:
            raise ValueError("bad modulation / rate")
        return c

    @classmethod
    def high_throughput(cls, modulation=QAM64):
        """presets::high_throughput(): pilots every 4th carrier (15 pilots / 44 data)."""
        return cls.default(modulation, use_pilots=1, pilot_spacing=4)

    def getSymbolDuration(self) -> int:
        return lib().ria_ofdm_symbol_samples(C.addressof(self))

    def getDataCarriers(self) -> int:
        return lib().ria_ofdm_data_carriers(C.addressof(self))

    def getPilotCarriers(self) -> int:
        return lib().ria_ofdm_pilot_carriers(C.addressof(self))

    def bitsPerSymbol(self) -> int:
        return self.getDataCarriers() * getBitsPerSymbol(self.modulation)


class FrameStatus(C.Structure):
    """ria_frame_status (include/ria_b200.h)."""
    _fields_ = [("cw_ok", C.c_uint8 * 4), ("cw_iters", C.c_int32 * 4), ("all_ok", C.c_uint8),
                ("header_valid", C.c_uint8), ("frame_crc_ok", C.c_uint8), ("type", C.c_uint8),
                ("seq", C.c_uint16), ("payload_len", C.c_uint16), ("src_hash", C.c_uint32),
                ("dst_hash", C.c_uint32), ("total_cw", C.c_uint8), ("ladder_cw_mask", C.c_uint8),
                ("ladder_max_attempt", C.c_uint8), ("fp_repair", C.c_uint8)]


FRAME_STATUS_DTYPE = np.dtype([
    ("cw_ok", np.uint8, 4), ("cw_iters", np.int32, 4), ("all_ok", np.uint8), ("header_valid", np.uint8),
    ("frame_crc_ok", np.uint8), ("type", np.uint8), ("seq", np.uint16), ("payload_len", np.uint16),
    ("src_hash", np.uint32), ("dst_hash", np.uint32), ("total_cw", np.uint8), ("ladder_cw_mask", np.uint8),
    ("ladder_max_attempt", np.uint8), ("fp_repair", np.uint8)],
    align=True)
assert FRAME_STATUS_DTYPE.itemsize == C.sizeof(FrameStatus)


def crc16(data: bytes) -> int:
    """ControlFrame::calculateCRC (src/protocol/frame_v2.cpp:115-128)."""
    buf = (C.c_uint8 * len(data)).from_buffer_copy(bytes(data))
    return int(lib().ria_crc16(C.addressof(buf), len(data)))


def channel_interleaver_step(bits_per_symbol: int, total_bits: int = 648) -> int:
    return int(lib().ria_channel_interleaver_step(int(bits_per_symbol), int(total_bits)))


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class OFDMDemodulator:
    """Presynced path of ultra::OFDMDemodulator, batched."""

    def __init__(self, config: ModemConfig, ctx: Optional[Context] = None):
        self.config = config
        self._ctx = ctx
        config.getSymbolDuration()

    @property
    def ctx(self) -> Context:
        if self._ctx is None:
            self._ctx = default_context()
        return self._ctx

    def soft_bits_per_frame(self, frame_len: int) -> int:
        n_sym = frame_len // self.config.getSymbolDuration()
        return max(0, n_sym - 2) * self.config.bitsPerSymbol()

    def process_presynced_batch(self, samples: torch.Tensor, cfo_hz: Optional[torch.Tensor] = None,
                                phase: Optional[torch.Tensor] = None, taps: bool = False):
        """samples: CUDA fp32 [n_frames, frame_len] (each row starts at the first LTS symbol).

        Returns dict(llr [n, stride], n_llr [n], snr_db, cfo, fading [, bins, h_lts])."""
        if not (isinstance(samples, torch.Tensor) and samples.is_cuda):
            raise RiaError("process_presynced_batch wants CUDA tensors (no CPU fallback)")
        if samples.dtype != torch.float32 or samples.dim() != 2:
            raise ValueError("samples must be fp32 [n_frames, frame_len]")
        if samples.stride(1) != 1:
            samples = samples.contiguous()
        n, frame_len = samples.shape
        dev = samples.device
        n_llr = self.soft_bits_per_frame(frame_len)
        stride = max(4, (n_llr + 3) & ~3)
        out = dict(
            llr=torch.empty((n, stride), dtype=torch.float32, device=dev),
            n_llr=torch.empty((n,), dtype=torch.int32, device=dev),
            snr_db=torch.empty((n,), dtype=torch.float32, device=dev),
            cfo=torch.empty((n,), dtype=torch.float32, device=dev),
            fading=torch.empty((n,), dtype=torch.float32, device=dev))
        nc = self.config.num_carriers
        n_sym = frame_len // self.config.getSymbolDuration()
        bins = h_lts = None
        if taps:
            bins = torch.zeros((n, max(n_sym, 1), nc, 2), dtype=torch.float32, device=dev)
            h_lts = torch.zeros((n, nc, 2), dtype=torch.float32, device=dev)
        for t in (cfo_hz, phase):
            if t is not None and not (t.is_cuda and t.dtype == torch.float32 and t.numel() == n):
                raise ValueError("cfo_hz / phase must be CUDA fp32 [n_frames]")
        ctx = self.ctx
        ctx.set_stream(torch.cuda.current_stream(dev))
        ctx.check(lib().ria_ofdm_presynced_batch_taps_dev(
            ctx.handle, C.addressof(self.config), _ptr(samples), samples.stride(0), frame_len,
            _ptr(cfo_hz), _ptr(phase), n, _ptr(out["llr"]), stride, _ptr(out["n_llr"]),
            _ptr(out["snr_db"]), _ptr(out["cfo"]), _ptr(out["fading"]), _ptr(bins), _ptr(h_lts)))
        if taps:
            out["bins"] = torch.view_as_complex(bins)
            out["h_lts"] = torch.view_as_complex(h_lts)
        return out


def decode_fixed_frame_batch(soft: torch.Tensor, rate: int, use_channel_interleave: bool,
                             bits_per_symbol: int, ctx: Optional[Context] = None, retry_ladder: Optional[bool] = None,
                             fp_repair: Optional[bool] = None):
    """v2::decodeFixedFrame for a batch: soft CUDA fp32 [n, >=2592].  First pass only, or with the
    LDPC retry ladder (frame_v2.cpp:1389-1546) when retry_ladder is True (None: the context's
    decode flags decide) and the false-positive repair (:1558-1916) when fp_repair is True; both = the
    complete reference function.

    Returns (data u8 [n, 4*bytes_per_cw], status structured array on the device as uint8 tensor
    viewable with FRAME_STATUS_DTYPE after .cpu().numpy())."""
    if not (isinstance(soft, torch.Tensor) and soft.is_cuda and soft.dtype == torch.float32 and soft.dim() == 2):
        raise RiaError("decode_fixed_frame_batch wants a CUDA fp32 [n, >=2592] tensor")
    if soft.stride(1) != 1:
        soft = soft.contiguous()
    ctx = ctx or default_context()
    n = soft.shape[0]
    bpc = code_params(rate)[0] // 8
    data = torch.empty((n, 4 * bpc), dtype=torch.uint8, device=soft.device)
    status = torch.empty((n, FRAME_STATUS_DTYPE.itemsize), dtype=torch.uint8, device=soft.device)
    ctx.set_stream(torch.cuda.current_stream(soft.device))
    saved = ctx.get_decode_flags()
    if retry_ladder is not None:
        ctx.set_decode_flags((saved | DECODE_RETRY_LADDER) if retry_ladder else (saved & ~DECODE_RETRY_LADDER))
    if fp_repair is not None:
        cur = ctx.get_decode_flags()
        ctx.set_decode_flags((cur | DECODE_FP_REPAIR) if fp_repair else (cur & ~DECODE_FP_REPAIR))
    try:
        ctx.check(lib().ria_frame_decode_batch_dev(
            ctx.handle, int(rate), int(bool(use_channel_interleave)), int(bits_per_symbol),
            _ptr(soft), soft.stride(0), n, _ptr(data), _ptr(status)))
    finally:
        ctx.set_decode_flags(saved)
    return data, status


def encode_fixed_frame_batch(frames: torch.Tensor, rate: int, use_channel_interleave: bool, bits_per_symbol: int,
                             ctx: Optional[Context] = None) -> torch.Tensor:
    """v2::encodeFixedFrame for a batch on the device: frames CUDA u8 [n, <= 4*bytes_per_cw] -> coded u8 [n, 324]
    (src/protocol/frame_v2.cpp:1285-1328)."""
    if not (isinstance(frames, torch.Tensor) and frames.is_cuda and frames.dtype == torch.uint8 and frames.dim() == 2):
        raise RiaError("encode_fixed_frame_batch wants a CUDA uint8 [n, frame_len] tensor")
    frames = frames.contiguous()
    ctx = ctx or default_context()
    n, flen = frames.shape
    coded = torch.empty((n, 324), dtype=torch.uint8, device=frames.device)
    ctx.set_stream(torch.cuda.current_stream(frames.device))
    ctx.check(lib().ria_encode_fixed_frame_batch_dev(ctx.handle, int(rate), int(bool(use_channel_interleave)),
                                                      int(bits_per_symbol), _ptr(frames), frames.stride(0), flen, n, _ptr(coded)))
    return coded


def ofdm_tx_frames(config: ModemConfig, coded: torch.Tensor, ctx: Optional[Context] = None) -> torch.Tensor:
    """OFDMModulator::generateTrainingSymbols + modulate for a batch on the device
    (src/ofdm/modulator.cpp:528-582, 348-477): coded CUDA u8 [n, coded_len] -> samples fp32 [n, frame_len],
    sample-identical to the reference transmitter."""
    if not (isinstance(coded, torch.Tensor) and coded.is_cuda and coded.dtype == torch.uint8 and coded.dim() == 2):
        raise RiaError("ofdm_tx_frames wants a CUDA uint8 [n, coded_len] tensor")
    coded = coded.contiguous()
    ctx = ctx or default_context()
    n, clen = coded.shape
    flen = lib().ria_ofdm_tx_frame_samples(C.addressof(config), clen)
    if flen <= 0:
        raise RiaError("ofdm_tx_frames: unsupported configuration")
    out = torch.empty((n, flen), dtype=torch.float32, device=coded.device)
    ctx.set_stream(torch.cuda.current_stream(coded.device))
    ctx.check(lib().ria_ofdm_tx_frames_dev(ctx.handle, C.addressof(config), _ptr(coded), coded.stride(0), clen, n,
                                           _ptr(out), out.stride(0)))
    return out


def ofdm_cox_tx_frames(config: ModemConfig, coded: torch.Tensor, ctx: Optional[Context] = None) -> torch.Tensor:
    """OFDMModulator::generatePreamble + modulate for a batch on the device (src/ofdm/modulator.cpp:479-532, 348-477) =
    what OFDMNvisWaveform (OFDM_COX) transmits: [silence][4 x STS][2 x LTS][data], sample-identical to the reference."""
    if not (isinstance(coded, torch.Tensor) and coded.is_cuda and coded.dtype == torch.uint8 and coded.dim() == 2):
        raise RiaError("ofdm_cox_tx_frames wants a CUDA uint8 [n, coded_len] tensor")
    coded = coded.contiguous()
    ctx = ctx or default_context()
    n, clen = coded.shape
    flen = lib().ria_ofdm_cox_tx_frame_samples(C.addressof(config), clen)
    if flen <= 0:
        raise RiaError("ofdm_cox_tx_frames: unsupported configuration")
    out = torch.empty((n, flen), dtype=torch.float32, device=coded.device)
    ctx.set_stream(torch.cuda.current_stream(coded.device))
    ctx.check(lib().ria_ofdm_cox_tx_frames_dev(ctx.handle, C.addressof(config), _ptr(coded), coded.stride(0), clen, n,
                                               _ptr(out), out.stride(0)))
    return out


def status_array(status: torch.Tensor) -> np.ndarray:
    return status.cpu().numpy().view(FRAME_STATUS_DTYPE).reshape(-1)


class OfdmRxChain:
    """Presynced OFDM data frames -> info bytes + per-frame status, one C call per batch."""

    def __init__(self, config: ModemConfig, rate: int, use_channel_interleave: bool = True,
                 ctx: Optional[Context] = None):
        self.config = config
        self.rate = int(rate)
        self.use_ci = bool(use_channel_interleave)
        self._ctx = ctx
        self.bytes_per_cw = code_params(rate)[0] // 8

    @property
    def ctx(self) -> Context:
        if self._ctx is None:
            self._ctx = default_context()
        return self._ctx

    def process_batch(self, samples: torch.Tensor, cfo_hz=None, phase=None):
        if not (isinstance(samples, torch.Tensor) and samples.is_cuda and samples.dtype == torch.float32):
            raise RiaError("process_batch wants CUDA fp32 [n_frames, frame_len] (no CPU fallback)")
        if samples.stride(1) != 1:
            samples = samples.contiguous()
        n, frame_len = samples.shape
        dev = samples.device
        data = torch.empty((n, 4 * self.bytes_per_cw), dtype=torch.uint8, device=dev)
        status = torch.empty((n, FRAME_STATUS_DTYPE.itemsize), dtype=torch.uint8, device=dev)
        snr = torch.empty((n,), dtype=torch.float32, device=dev)
        ctx = self.ctx
        ctx.set_stream(torch.cuda.current_stream(dev))
        ctx.check(lib().ria_ofdm_rx_frames_dev(
            ctx.handle, C.addressof(self.config), self.rate, int(self.use_ci), _ptr(samples),
            samples.stride(0), frame_len, _ptr(cfo_hz), _ptr(phase), n, _ptr(data), _ptr(status), _ptr(snr)))
        return data, status, snr

    def process_batch_host(self, samples: np.ndarray, cfo_hz=None, phase=None):
        """Host buffers (numpy, ideally backed by pinned memory) through ria_ofdm_rx_frames_host."""
        assert samples.dtype == np.float32 and samples.ndim == 2 and samples.strides[1] == 4
        n, frame_len = samples.shape
        data = np.empty((n, 4 * self.bytes_per_cw), np.uint8)
        status = np.empty(n, FRAME_STATUS_DTYPE)
        snr = np.empty(n, np.float32)
        ctx = self.ctx
        ctx.check(lib().ria_ofdm_rx_frames_host(
            ctx.handle, C.addressof(self.config), self.rate, int(self.use_ci),
            samples.ctypes.data, samples.strides[0] // 4, frame_len,
            cfo_hz.ctypes.data if cfo_hz is not None else None,
            phase.ctypes.data if phase is not None else None,
            n, data.ctypes.data, status.ctypes.data, snr.ctypes.data))
        return data, status, snr


class OfdmCoxRxChain(OfdmRxChain):
    """OFDM_COX windows -> info bytes + per-frame status, one C call per batch (``ria_ofdm_cox_rx_frames_dev`` / ``_host``):
    OFDMNvisWaveform::detectSync (Schmidl-Cox search) -> process at the LTS position with the CFO and phase found -> frame
    decode.  ``frame_len`` = samples from the first LTS symbol to the end of the frame."""

    def process_windows(self, windows: torch.Tensor, frame_len: int, threshold: float = 0.8, noise_floor=None):
        """windows CUDA fp32 [n, window] -> (data, status, snr, sync uint8 [n, 32])"""
        from .sync import SYNC_RESULT_DTYPE
        if not (isinstance(windows, torch.Tensor) and windows.is_cuda and windows.dtype == torch.float32 and windows.dim() == 2):
            raise RiaError("process_windows wants CUDA fp32 [n, window] (no CPU fallback)")
        if windows.stride(1) != 1:
            windows = windows.contiguous()
        n, window = windows.shape
        dev = windows.device
        data = torch.empty((n, 4 * self.bytes_per_cw), dtype=torch.uint8, device=dev)
        status = torch.empty((n, FRAME_STATUS_DTYPE.itemsize), dtype=torch.uint8, device=dev)
        snr = torch.empty((n,), dtype=torch.float32, device=dev)
        sync = torch.empty((n, SYNC_RESULT_DTYPE.itemsize), dtype=torch.uint8, device=dev)
        ctx = self.ctx
        ctx.set_stream(torch.cuda.current_stream(dev))
        ctx.check(lib().ria_ofdm_cox_rx_frames_dev(
            ctx.handle, C.addressof(self.config), self.rate, int(self.use_ci), _ptr(windows), windows.stride(0), window,
            int(frame_len), float(threshold), _ptr(noise_floor), n, _ptr(data), _ptr(status), _ptr(snr), _ptr(sync)))
        return data, status, snr, sync

    def process_windows_host(self, windows: np.ndarray, frame_len: int, threshold: float = 0.8):
        """Host windows (numpy, ideally pinned) through ria_ofdm_cox_rx_frames_host -> (data, status, snr, sync)"""
        from .sync import SYNC_RESULT_DTYPE
        assert windows.dtype == np.float32 and windows.ndim == 2 and windows.strides[1] == 4
        n, window = windows.shape
        data = np.empty((n, 4 * self.bytes_per_cw), np.uint8)
        status = np.empty(n, FRAME_STATUS_DTYPE)
        snr = np.empty(n, np.float32)
        sync = np.empty(n, SYNC_RESULT_DTYPE)
        ctx = self.ctx
        ctx.check(lib().ria_ofdm_cox_rx_frames_host(
            ctx.handle, C.addressof(self.config), self.rate, int(self.use_ci), windows.ctypes.data, windows.strides[0] // 4,
            window, int(frame_len), float(threshold), None, n, data.ctypes.data, status.ctypes.data, snr.ctypes.data,
            sync.ctypes.data))
        return data, status, snr, sync


class OFDMChirpWaveform:
    """RX half of ultra::OFDMChirpWaveform (IWaveform) with batch = 1 semantics.

    configure / setFrequencyOffset / setAbsoluteTrainingPosition / process / getSoftBits /
    estimatedSNR / estimatedCFO / getFadingIndex / reset behave like the reference class
    (src/waveform/ofdm_chirp_waveform.cpp); detectSync/detectDataSync are provided by
    ria_b200.sync once the correlator kernels land."""

    def __init__(self, config: Optional[ModemConfig] = None, ctx: Optional[Context] = None):
        self._ctx = ctx
        if config is None:
            # default constructor: 512 FFT / 30 carriers is not a size the batched kernels are
            # built for; the 1024-FFT profile every tool uses is the default here
            config = ModemConfig.for_waveform(DQPSK, 2)
        self.config_ = config
        self.cfo_hz_ = 0.0
        self.last_cfo_ = 0.0
        self.last_snr_ = 0.0
        self.last_fading_ = 0.0
        self.training_start_sample_ = 0
        self.abs_training_start_ = None
        self.soft_bits_ = np.zeros(0, np.float32)
        self._dem = OFDMDemodulator(self.config_, ctx)

    def configure(self, mod: int, rate: int) -> None:
        if mod not in (DBPSK, DQPSK, D8PSK, QPSK, BPSK, QAM16, QAM32, QAM64):
            mod = DQPSK                                   # ofdm_chirp_waveform.cpp:81-87
        self.config_ = ModemConfig.for_waveform(mod, rate)
        self._dem = OFDMDemodulator(self.config_, self._ctx)

    def setFrequencyOffset(self, cfo_hz: float) -> None:
        self.cfo_hz_ = float(np.float32(cfo_hz))

    def setAbsoluteTrainingPosition(self, pos: int) -> None:
        self.abs_training_start_ = int(pos)

    def getSamplesPerSymbol(self) -> int:
        return self.config_.getSymbolDuration()

    def process(self, samples) -> bool:
        samples = np.ascontiguousarray(samples, dtype=np.float32)
        if samples.size < self.getSamplesPerSymbol():
            return False
        # initial CFO phase, ofdm_chirp_waveform.cpp:404-413 (fp32 with double-promoted pi)
        ref = self.abs_training_start_ if self.abs_training_start_ is not None else self.training_start_sample_
        ph = np.float32(-2.0 * np.pi * np.float64(np.float32(self.cfo_hz_)) * ref / self.config_.sample_rate)
        # the reference evaluates -2.0f*M_PI*cfo*pos/sample_rate left to right in double
        while float(ph) > np.pi:
            ph = np.float32(np.float64(ph) - 2.0 * np.pi)
        while float(ph) < -np.pi:
            ph = np.float32(np.float64(ph) + 2.0 * np.pi)
        dev = torch.device("cuda", self._dem.ctx.device)
        x = torch.from_numpy(samples).to(dev).unsqueeze(0)
        cfo = torch.tensor([self.cfo_hz_], dtype=torch.float32, device=dev)
        pht = torch.tensor([float(ph)], dtype=torch.float32, device=dev)
        out = self._dem.process_presynced_batch(x, cfo, pht)
        torch.cuda.synchronize(dev)
        n = int(out["n_llr"][0].item())
        ready = n >= 648
        if ready:
            self.soft_bits_ = out["llr"][0, :n].cpu().numpy()
            self.last_snr_ = float(out["snr_db"][0].item())
            self.cfo_hz_ = float(out["cfo"][0].item())
            self.last_cfo_ = self.cfo_hz_
        self.last_fading_ = float(out["fading"][0].item())
        return ready

    def getSoftBits(self) -> np.ndarray:
        bits, self.soft_bits_ = self.soft_bits_, np.zeros(0, np.float32)   # moves out (:470-472)
        return bits

    def estimatedSNR(self) -> float:
        return self.last_snr_

    def estimatedCFO(self) -> float:
        return self.last_cfo_ if abs(self.last_cfo_) > 0.1 else self.cfo_hz_

    def getFadingIndex(self) -> float:
        return self.last_fading_

    def reset(self) -> None:
        self.soft_bits_ = np.zeros(0, np.float32)
        self.abs_training_start_ = None               # CFO is preserved across reset (:474-485)


# ---------------------------------------------------------------------------------------------
# Frame-level decode of OFDM receptions: StreamingDecoder::decodeFrame for a batch
# ---------------------------------------------------------------------------------------------
class OfdmFrameDecoder:
    """StreamingDecoder::decodeFrame (src/gui/modem/streaming_decoder.cpp:2820-3056) for a batch of OFDM receptions, the
    reference's "try both" strategy:
      1. R1/4 control fast path when the data rate is not R1/4 (:2864-2886): codec decode of the first 648 soft bits
         at R1/4; a valid one-codeword header ends the frame;
      2. codeword 0 decoded raw at the data rate (:2892-2925): a one-codeword header ends the frame, a four-codeword
         header (or no header at all) goes on to the frame-interleaved decode;
      3. v2::decodeFixedFrame -- complete, with retry ladder and false-positive repair -- and reassembly
         (:2928-2967); when it fails, the one-codeword salvage with robustDecodeSingleCW at R1/4, then at the data
         rate (:2972-3008);
      4. the legacy sequential path for headers that announce another codeword count (:3013-3053).
    Every LDPC decode runs on the device, batched over the receptions that reach the step; the branching between the
    steps is host logic, as in the reference.  rate = the connected data rate (a disconnected receiver decodes R1/4)."""

    def __init__(self, modulation: int, rate: int, data_carriers: int, connected: bool = True,
                 use_channel_interleave: bool = True, ctx: Optional[Context] = None):
        from . import fec
        self.modulation, self.data_carriers = int(modulation), int(data_carriers)
        self.rate = int(rate) if connected else 0                     # :2831
        self.use_ci = bool(use_channel_interleave)
        self._ctx = ctx
        self.k = fec.code_params(self.rate)[0]
        self.bpc = self.k // 8
        self.bps = self.data_carriers * getBitsPerSymbol(self.modulation)

        def codec(r):                                                 # LDPCCodec::decode: factor 0.75, recommended iterations
            d = fec.LDPCDecoder(r, ctx)
            d.setMaxIterations(fec.recommended_iterations(r))
            d.setMinSumFactor(0.75)
            return d
        self.codec = codec(self.rate)
        self.codec_r14 = codec(0) if self.rate != 0 else None
        self.robust = fec.LDPCDecoder(self.rate, ctx)
        self.robust_r14 = fec.LDPCDecoder(0, ctx)

    @staticmethod
    def _header(d0, d0len):
        """v2::parseHeader on rows of codeword-0 bytes -> (valid, type, total_cw, payload_len, is_control)"""
        from .mcdpsk import _CONTROL_TYPES, _crc16_rows
        ftype = d0[:, 2].astype(np.int32)
        is_control = np.isin(ftype, _CONTROL_TYPES)
        crc_ctl = _crc16_rows(d0[:, :18]) == ((d0[:, 18].astype(np.uint16) << 8) | d0[:, 19])
        crc_dat = _crc16_rows(d0[:, :15]) == ((d0[:, 15].astype(np.uint16) << 8) | d0[:, 16])
        valid = (d0len >= 20) & (d0[:, 0] == 0x55) & (d0[:, 1] == 0x4C) & np.where(is_control, crc_ctl, crc_dat)
        total_cw = np.where(is_control, 1, d0[:, 12].astype(np.int32))
        payload_len = np.where(is_control, 0, (d0[:, 13].astype(np.int32) << 8) | d0[:, 14])
        return valid, ftype, total_cw, payload_len, is_control

    def _decode_cw(self, dec, rows: torch.Tensor, robust: bool):
        stride = 64                                                   # ceil(486 / 8) = 61 bytes at R3/4
        if robust:
            info, ok, _, _ = dec.robust_decode_batch(rows.contiguous(), info_stride=stride)
        else:
            info, ok, _ = dec.decode_batch(rows.contiguous(), info_stride=stride)
        return info.cpu().numpy(), ok.cpu().numpy().astype(bool)

    def decode_batch(self, soft: torch.Tensor):
        """soft: CUDA fp32 [n, L], L >= 648 soft bits per reception.  Returns numpy arrays success u8[n], frame_type i32[n],
        codewords_ok / codewords_failed i32[n], frame_len i32[n], frame u8[n, W]."""
        from .fec import code_params
        if not (isinstance(soft, torch.Tensor) and soft.is_cuda and soft.dtype == torch.float32 and soft.dim() == 2):
            raise RiaError("decode_batch wants CUDA fp32 [n, L] (no CPU fallback)")
        n, L = soft.shape
        bpc = self.bpc
        W = max(4 * bpc, (L // 648) * max(bpc, 21), 64)
        out = dict(success=np.zeros(n, np.uint8), frame_type=np.full(n, 0x10, np.int32), codewords_ok=np.zeros(n, np.int32),
                   codewords_failed=np.zeros(n, np.int32), frame_len=np.zeros(n, np.int32), frame=np.zeros((n, W), np.uint8))
        if n == 0 or L < 648:
            return out
        dev = soft.device
        alive = np.ones(n, bool)                                      # receptions that have not returned yet

        def finish_control(idx, data, dlen, ftype):
            out["success"][idx] = 1
            out["codewords_ok"][idx] = 1
            out["codewords_failed"][idx] = 0
            out["frame_type"][idx] = ftype
            out["frame"][idx, :64] = 0
            for j, g in enumerate(idx):
                out["frame"][g, : dlen[j]] = data[j, : dlen[j]]
            out["frame_len"][idx] = dlen
            alive[idx] = False

        cw0 = soft[:, :648]
        # ---- 1. R1/4 control fast path ----
        if self.codec_r14 is not None:
            info, ok = self._decode_cw(self.codec_r14, cw0, robust=False)
            dlen = np.full(n, 20, np.int32)                           # data_r14.resize(bytes per codeword of R1/4)
            d = info.copy(); d[:, 20:] = 0
            valid, ftype, total_cw, _, _ = self._header(d, dlen)
            hit = np.nonzero(ok & valid & (total_cw == 1))[0]
            if len(hit):
                finish_control(hit, d[hit], dlen[hit], ftype[hit])
        # ---- 2. codeword 0 raw at the data rate ----
        info0, ok0 = self._decode_cw(self.codec, cw0, robust=False)
        d0 = info0.copy(); d0[:, bpc:] = 0
        d0len = np.full(n, bpc, np.int32)
        magic0 = ok0 & (d0[:, 0] == 0x55) & (d0[:, 1] == 0x4C)
        valid0, ftype0, total0, payload0, is_ctl0 = self._header(d0, d0len)
        valid0 &= magic0
        out["frame_type"][alive & valid0] = ftype0[alive & valid0]
        hit = np.nonzero(alive & valid0 & (total0 == 1))[0]
        if len(hit):
            finish_control(hit, d0[hit], d0len[hit], ftype0[hit])
        try_fi = alive & (~magic0 | (valid0 & (total0 == 4)))
        # ---- 3. frame-interleaved decode of the four-codeword frame ----
        fi = np.nonzero(try_fi & (L >= 2592))[0] if L >= 2592 else np.zeros(0, np.int64)
        if len(fi):
            sel = soft.index_select(0, torch.from_numpy(fi).to(dev))
            data, status = decode_fixed_frame_batch(sel, self.rate, self.use_ci, self.bps, self._ctx,
                                                    retry_ladder=True, fp_repair=True)
            data = data.cpu().numpy()
            st = status_array(status)
            n_ok = (st["cw_ok"] != 0).sum(axis=1)
            out["codewords_ok"][fi] = n_ok
            out["codewords_failed"][fi] = 4 - n_ok
            allok = n_ok == 4
            # CodewordStatus::reassemble: header of codeword 0, then chunks 1..3 (a chunk that starts with 0xD5 loses two bytes)
            for j in np.nonzero(allok)[0]:
                g = fi[j]
                chunks = data[j].reshape(4, bpc)
                hv, _, _, plen, isc = self._header(chunks[:1], np.array([bpc]))
                frame = b""
                if hv[0]:
                    expected = 20 if isc[0] else 17 + int(plen[0]) + 2
                    parts, have = [], 0
                    for c in range(4):
                        if have >= expected:
                            break
                        ch = chunks[c]
                        if c > 0 and bpc >= 2 and ch[0] == 0xD5:
                            ch = ch[2:]
                        ch = ch[: expected - have]
                        parts.append(ch); have += len(ch)
                    frame = np.concatenate(parts).tobytes()
                out["success"][g] = 1 if frame else 0                # 4/4 but reassemble failed: LDPC false positive (:2946-2950)
                out["frame"][g, : len(frame)] = np.frombuffer(frame, np.uint8)
                out["frame_len"][g] = len(frame)
                if len(frame) >= 3:
                    out["frame_type"][g] = frame[2]
                alive[g] = False
            # one-codeword salvage of what the four-codeword decode could not read (:2972-3008)
            bad = fi[~allok]
            for dec, r in ((self.robust_r14, 0), (self.robust, self.rate)):
                if len(bad) == 0 or (r == self.rate and self.rate == 0 and dec is self.robust):
                    continue
                infoS, okS = self._decode_cw(dec, cw0.index_select(0, torch.from_numpy(bad).to(dev)), robust=True)
                bpc_s = code_params(r)[0] // 8
                dS = infoS.copy(); dS[:, bpc_s:] = 0
                lenS = np.full(len(bad), bpc_s, np.int32)
                vS, tS, totS, _, _ = self._header(dS, lenS)
                hitS = okS & vS & (totS == 1)
                if hitS.any():
                    finish_control(bad[hitS], dS[hitS], lenS[hitS], tS[hitS])
                bad = bad[~hitS]
        # ---- 4. legacy sequential path (:3013-3053) ----
        leg = np.nonzero(alive & magic0)[0]
        if len(leg):
            out["codewords_ok"][leg] = 1
            out["codewords_failed"][leg] = np.where(np.isin(leg, fi), out["codewords_failed"][leg], 0)
            keep = valid0[leg]
            leg = leg[keep]
            avail = L // 648
            short = total0[leg] > avail
            out["frame"][leg[short], :64] = d0[leg[short]]
            out["frame_len"][leg[short]] = bpc
            leg = leg[~short]
            if len(leg):
                tmax = int(total0[leg].max())
                ci = None
                if self.use_ci:                                       # ChannelInterleaver::deinterleave: out[i] = in[(i * step) % 648]
                    step = channel_interleaver_step(self.bps, 648)
                    ci = torch.from_numpy((np.arange(648) * step) % 648).to(dev)
                okm = np.zeros((len(leg), tmax), bool); okm[:, 0] = True
                datam = np.zeros((len(leg), tmax, bpc), np.uint8); datam[:, 0] = d0[leg, :bpc]
                rows_t = torch.from_numpy(leg).to(dev)
                for c in range(1, tmax):
                    need = np.nonzero(total0[leg] > c)[0]
                    if len(need) == 0:
                        break
                    bits = soft.index_select(0, rows_t[torch.from_numpy(need).to(dev)])[:, c * 648:(c + 1) * 648]
                    if ci is not None:
                        bits = bits.index_select(1, ci)
                    infoC, okC = self._decode_cw(self.codec, bits, robust=False)
                    okm[need, c] = okC
                    datam[need, c] = infoC[:, :bpc]
                for j, g in enumerate(leg):
                    t = int(total0[g])
                    nok = int(okm[j, :t].sum())
                    out["codewords_ok"][g] = nok
                    out["codewords_failed"][g] += t - nok              # on top of what a failed four-codeword decode counted
                    if nok == t:
                        expected = 20 if is_ctl0[g] else 17 + int(payload0[g]) + 2
                        parts, have = [], 0
                        for c in range(t):
                            if have >= expected:
                                break
                            ch = datam[j, c]
                            if c > 0 and bpc >= 2 and ch[0] == 0xD5:
                                ch = ch[2:]
                            ch = ch[: expected - have]
                            parts.append(ch); have += len(ch)
                        fr = np.concatenate(parts)
                        out["success"][g] = 1
                        out["frame"][g, : len(fr)] = fr
                        out["frame_len"][g] = len(fr)
        return out

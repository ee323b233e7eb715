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

    def process_batch_zc(self, rows: torch.Tensor, frame_len: int, sync_window: int, acc: torch.Tensor,
                         first_reception: bool = False, out=None, known_cfo_hz: Optional[torch.Tensor] = None,
                         threshold: float = 0.2):
        """The same chain for connected-mode receptions behind the Zadoff-Chu data preamble
        (``ria_mcdpsk_zc_rx_frames_dev``): ZC detect (roots DATA | CONTROL) -> process -> chase -> LDPC."""
        from .sync import SYNC_RESULT_DTYPE, ZCConfig, ZC_ROOT_MASK_CONTROL, ZC_ROOT_MASK_DATA
        if not (isinstance(rows, torch.Tensor) and rows.is_cuda and rows.dtype == torch.float32 and rows.dim() == 2):
            raise RiaError("process_batch_zc wants CUDA fp32 [n_frames, row_len] (no CPU fallback)")
        if rows.stride(1) != 1:
            rows = rows.contiguous()
        n = rows.shape[0]
        dev = rows.device
        info_stride = (self.info_bytes + 3) & ~3
        if out is None:
            out = dict(info=torch.empty((n, info_stride), dtype=torch.uint8, device=dev),
                       ok=torch.empty((n,), dtype=torch.uint8, device=dev),
                       iters=torch.empty((n,), dtype=torch.int32, device=dev),
                       sync=torch.empty((n, SYNC_RESULT_DTYPE.itemsize), dtype=torch.uint8, device=dev))
        out["acc"] = acc
        if not hasattr(self, "_zc"):
            self._zc = ZCConfig.default()
        ctx = self.ctx
        ctx.set_stream(torch.cuda.current_stream(dev))
        ctx.check(lib().ria_mcdpsk_zc_rx_frames_dev(
            ctx.handle, C.addressof(self.config), C.addressof(self._zc), _ptr(rows), rows.stride(0), int(sync_window),
            int(frame_len), _ptr(known_cfo_hz), float(threshold), ZC_ROOT_MASK_DATA | ZC_ROOT_MASK_CONTROL, n, self.rate,
            self.max_iter, self.factor, _ptr(acc), int(bool(first_reception)), _ptr(out["info"]), out["info"].stride(0),
            _ptr(out["ok"]), _ptr(out["iters"]), _ptr(out["sync"])))
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


# ---------------------------------------------------------------------------------------------
# Frame-level decode of MC-DPSK receptions: StreamingDecoder::decodeMCDPSKFrame for a batch
# ---------------------------------------------------------------------------------------------
_CONTROL_TYPES = (0x10, 0x11, 0x16, 0x17, 0x20, 0x21, 0x15, 0x40)      # v2::isControlFrame (frame_v2.hpp:222-228)
_CONNECT_TYPES = (0x12, 0x13, 0x14)                                     # isConnectFrame minus DISCONNECT (:348-351)
_CRC_TABLE = None


def _crc16_rows(rows):
    """CRC-16/CCITT-FALSE (ControlFrame::calculateCRC, frame_v2.cpp:115-128) of every row of a uint8 [n, L] array."""
    import numpy as np
    global _CRC_TABLE
    if _CRC_TABLE is None:
        t = np.zeros(256, np.uint16)
        for b in range(256):
            r = b << 8
            for _ in range(8):
                r = ((r << 1) ^ 0x1021) & 0xFFFF if r & 0x8000 else (r << 1) & 0xFFFF
            t[b] = r
        _CRC_TABLE = t
    crc = np.full(rows.shape[0], 0xFFFF, np.uint16)
    for j in range(rows.shape[1]):
        crc = ((crc << 8) & 0xFFFF) ^ _CRC_TABLE[((crc >> 8) ^ rows[:, j]) & 0xFF]
    return crc


class McdpskFrameDecoder:
    """StreamingDecoder::decodeMCDPSKFrame (src/gui/modem/streaming_decoder.cpp:2580-2822) for a batch of receptions.

    Per reception: robustDecodeSingleCW on codeword 0 (with the R1/4 fallback for control frames, :2639-2649),
    v2::parseHeader, the CONNECT sanity rule (:2714-2723); a 1-codeword frame is done; a multi-codeword frame
    decodes codewords 1.. with robustDecodeSingleCW, failed ones go through the HARQ chase cache (store, and when
    earlier receptions are cached, decode the combined soft bits, :2762-2789); CodewordStatus::reassemble builds
    the frame.  Every LDPC decode and every soft-bit combine of the batch runs on the device (one robust-decode
    launch for all codewords, one for the R1/4 fallbacks, one combine + one decode for the HARQ retries); the
    frame logic around them is host logic, as in the reference.  The MC-DPSK channel interleaver is off
    (StreamingDecoder's default, streaming_decoder.hpp:373).

    Receptions of a batch are applied to the cache in order, like consecutive decodeMCDPSKFrame calls, except that
    all stores of the batch precede its markDecoded / removeEntry updates (only observable when the cache evicts)."""

    def __init__(self, rate: int, ctx: Optional[Context] = None, chase_cache=None):
        from . import fec
        self.rate = int(rate)
        self._ctx = ctx
        self.k = fec.code_params(self.rate)[0]
        self.bpc = self.k // 8                                 # v2::getBytesPerCodeword
        self.dec = fec.LDPCDecoder(self.rate, ctx)
        self.dec_r14 = fec.LDPCDecoder(0, ctx) if self.rate != 0 else None
        self.cache = chase_cache
        self.stats = dict(chase_stores=0, chase_recoveries=0, r14_fallbacks=0)

    def decode_batch(self, soft: torch.Tensor):
        """soft: CUDA fp32 [n, C * 648] (C codeword slots per reception).  Returns a dict of numpy arrays:
        success u8[n], frame_type i32[n], codewords_ok i32[n], codewords_failed i32[n], frame_len i32[n],
        frame u8[n, C * bytes_per_cw] (frame_len valid bytes each)."""
        import numpy as np
        if not (isinstance(soft, torch.Tensor) and soft.is_cuda and soft.dtype == torch.float32 and soft.dim() == 2):
            raise RiaError("decode_batch wants CUDA fp32 [n, C * 648] (no CPU fallback)")
        n, width = soft.shape
        C_slots = width // 648
        bpc = self.bpc
        out = dict(success=np.zeros(n, np.uint8), frame_type=np.full(n, 0x10, np.int32), codewords_ok=np.zeros(n, np.int32),
                   codewords_failed=np.zeros(n, np.int32), frame_len=np.zeros(n, np.int32),
                   frame=np.zeros((n, max(1, C_slots) * max(bpc, 20)), np.uint8))
        if n == 0 or C_slots == 0:
            return out
        cw = soft[:, : C_slots * 648].contiguous().view(n * C_slots, 648)
        stride = (((self.k + 7) // 8) + 3) & ~3
        info_d, ok_d, _, _ = self.dec.robust_decode_batch(cw, info_stride=stride)
        info = info_d.cpu().numpy().reshape(n, C_slots, stride)
        ok = ok_d.cpu().numpy().reshape(n, C_slots).astype(bool)
        d0 = np.zeros((n, max(stride, 24)), np.uint8)
        d0[:, :stride] = info[:, 0, :]
        d0len = np.full(n, bpc, np.int32)                       # data0.resize(bytes_per_cw) (:2686)
        good0 = ok[:, 0] & (d0[:, 0] == 0x55) & (d0[:, 1] == 0x4C)
        if self.dec_r14 is not None and (~good0).any():       # control frames are always R1/4 (:2639-2649)
            idx = np.nonzero(~good0)[0]
            sel = soft[torch.from_numpy(idx).to(soft.device), :648].contiguous()
            i2, o2, _, _ = self.dec_r14.robust_decode_batch(sel, info_stride=24)
            i2, o2 = i2.cpu().numpy(), o2.cpu().numpy().astype(bool)
            hit = o2 & (i2[:, 0] == 0x55) & (i2[:, 1] == 0x4C)
            d0[idx[hit]] = 0
            d0[idx[hit], :24] = i2[hit]
            d0len[idx[hit]] = min(21, bpc)                     # the R1/4 decode returns ceil(162 / 8) bytes
            good0[idx[hit]] = True
            self.stats["r14_fallbacks"] += int(hit.sum())
        # ---- v2::parseHeader on the first bytes_per_cw bytes (frame_v2.cpp:1195-1253; needs >= 20 bytes) ----
        d0[np.arange(d0.shape[1])[None, :] >= d0len[:, None]] = 0
        ftype = d0[:, 2].astype(np.int32)
        is_control = np.isin(ftype, _CONTROL_TYPES)
        crc_ctl = _crc16_rows(d0[:, :18]) == ((d0[:, 18].astype(np.uint16) << 8) | d0[:, 19])
        crc_dat = _crc16_rows(d0[:, :15]) == ((d0[:, 15].astype(np.uint16) << 8) | d0[:, 16])
        valid = good0 & (d0len >= 20) & np.where(is_control, crc_ctl, crc_dat)
        total_cw = np.where(is_control, 1, d0[:, 12].astype(np.int32))
        payload_len = np.where(is_control, 0, (d0[:, 13].astype(np.int32) << 8) | d0[:, 14])
        seq = (d0[:, 4].astype(np.int32) << 8) | d0[:, 5]
        src = (d0[:, 6].astype(np.int32) << 16) | (d0[:, 7].astype(np.int32) << 8) | d0[:, 8]
        dst = (d0[:, 9].astype(np.int32) << 16) | (d0[:, 10].astype(np.int32) << 8) | d0[:, 11]
        # impossible CONNECT headers are decode false positives (:2714-2723)
        min_connect = max(2, ((17 + 25 + 2) * 8 + self.k - 1) // self.k)
        valid &= ~(np.isin(ftype, _CONNECT_TYPES) & (total_cw < min_connect))
        out["frame_type"][valid] = ftype[valid]
        out["codewords_ok"][valid] = 1
        # ---- one-codeword (control) frames ----
        one = valid & (total_cw == 1)
        W = out["frame"].shape[1]
        out["success"][one] = 1
        out["frame"][one, : min(W, d0.shape[1])] = d0[one, : min(W, d0.shape[1])]
        out["frame_len"][one] = d0len[one]
        # ---- multi-codeword frames ----
        multi = valid & (total_cw > 1)
        partial = multi & (total_cw > C_slots)                   # not all codewords are in the buffer yet (:2745-2751)
        out["frame"][partial, : min(W, d0.shape[1])] = d0[partial, : min(W, d0.shape[1])]
        out["frame_len"][partial] = d0len[partial]
        full = np.nonzero(multi & ~partial)[0]
        if len(full) == 0:
            return out
        cw_idx = np.arange(C_slots)[None, :]
        needed = (cw_idx >= 1) & (cw_idx < total_cw[full, None])             # codewords 1 .. total_cw - 1
        cw_ok = ok[full] & needed
        cw_data = info[full][:, :, :bpc].copy()
        cw0_full = d0[full]
        failed_f, failed_c = np.nonzero(needed & ~cw_ok)
        cache = self.cache
        if cache is not None and cache.enabled and len(failed_f):
            keys = [(int(seq[full[f]]), int(src[full[f]]), int(dst[full[f]])) for f in failed_f]
            rows = torch.from_numpy(full[failed_f] * C_slots + failed_c).to(soft.device)
            stored = cache.store_batch(keys, [int(c) for c in failed_c], [int(total_cw[full[f]]) for f in failed_f],
                                       cw.index_select(0, rows))
            self.stats["chase_stores"] += int(sum(stored))
            # decode the combined soft bits where earlier receptions are cached (:2771-2787); getCombined refuses
            # codewords the cache already holds as decoded
            retry, rows_c = [], []
            for j in range(len(keys)):
                if cache.getCombineCount(keys[j], int(failed_c[j])) > 1:
                    t = cache.getCombined(keys[j], int(failed_c[j]))
                    if t is not None:
                        retry.append(j)
                        rows_c.append(t)
            if retry:
                comb = torch.stack(rows_c)
                i3, o3, _, _ = self.dec.robust_decode_batch(comb.contiguous(), info_stride=stride)
                i3, o3 = i3.cpu().numpy(), o3.cpu().numpy().astype(bool)
                for t, j in enumerate(retry):
                    if o3[t]:
                        cw_ok[failed_f[j], failed_c[j]] = True
                        cw_data[failed_f[j], failed_c[j], :] = i3[t, :bpc]
                        cache.markDecoded(keys[j], int(failed_c[j]))
                        cache.stats["recoveries"] += 1
                        self.stats["chase_recoveries"] += 1
        n_ok = cw_ok.sum(axis=1)
        n_need = needed.sum(axis=1)
        out["codewords_ok"][full] = 1 + n_ok
        out["codewords_failed"][full] = n_need - n_ok
        all_ok = n_ok == n_need
        if cache is not None and cache.enabled and cache.entries:
            for f in range(len(full)):                            # decoded codewords leave the cache (:2797-2800, :2807-2810)
                key = (int(seq[full[f]]), int(src[full[f]]), int(dst[full[f]]))
                if key in cache.entries:
                    if all_ok[f]:
                        cache.removeEntry(key)
                    else:
                        for c in np.nonzero(cw_ok[f])[0]:
                            cache.markDecoded(key, int(c))
        # ---- CodewordStatus::reassemble (frame_v2.cpp:1030-1066, reassembleCodewords :960-985) ----
        done = np.nonzero(all_ok)[0]
        for f in done:
            g = full[f]
            expected = 20 if is_control[g] else 17 + int(payload_len[g]) + 2
            parts, have = [], 0
            for c in range(int(total_cw[g])):
                if have >= expected:
                    break
                chunk = cw_data[f, c] if c > 0 else cw0_full[f, : d0len[g]]
                if c > 0 and bpc >= 2 and chunk[0] == 0xD5:
                    chunk = chunk[2:]
                chunk = chunk[: expected - have]
                parts.append(chunk)
                have += len(chunk)
            fr = np.concatenate(parts)
            out["frame"][g, : len(fr)] = fr
            out["frame_len"][g] = len(fr)
            out["success"][g] = 1
        return out


class McdpskZcRxChain:
    """Connected-mode MC-DPSK receptions behind the Zadoff-Chu data preamble (SURVEY.md 8d, configs[2] variant ii), the
    calls a StreamingDecoder makes for such a frame: MCDPSKWaveform::detectDataSync (ZC, roots DATA | CONTROL, in the
    31 120-sample search window; src/waveform/mc_dpsk_waveform.cpp:227-292, streaming_decoder.cpp:423-435) ->
    MCDPSKWaveform::process at the detected training start with (known + residual) CFO (:294-338) ->
    decodeMCDPSKFrame (multi-codeword frames, HARQ chase combining; streaming_decoder.cpp:2580-2822).
    Synchronisation, demodulation, every LDPC decode and the soft-bit combining run on the device."""

    def __init__(self, config: MultiCarrierDPSKConfig, rate: int, ctx: Optional[Context] = None, chase_cache=None,
                 threshold: float = 0.2):
        from . import sync
        self.config, self.rate, self.threshold = config, int(rate), float(threshold)
        self._ctx = ctx
        self.zc = sync.ZCSync(None, ctx)
        self.dem = MCDPSKDemodulator(config, ctx)
        self.decoder = McdpskFrameDecoder(rate, ctx, chase_cache)

    def process_batch(self, rows: torch.Tensor, window: int, frame_len: int, known_cfo_hz: Optional[torch.Tensor] = None):
        """rows: CUDA fp32 [n, row_len >= window].  frame_len = samples handed to process() from the training start.
        Returns (decode dict of McdpskFrameDecoder.decode_batch, sync uint8 tensor [n, 32])."""
        from . import sync
        if not (isinstance(rows, torch.Tensor) and rows.is_cuda and rows.dtype == torch.float32 and rows.dim() == 2):
            raise RiaError("process_batch wants CUDA fp32 [n, row_len] (no CPU fallback)")
        n, row_len = rows.shape
        dev = rows.device
        res = self.zc.detect_batch(rows[:, :window], self.threshold, sync.ZC_ROOT_MASK_DATA | sync.ZC_ROOT_MASK_CONTROL,
                                   known_cfo_hz)
        f = res.view(torch.int32)                       # detected, start_sample, correlation, cfo_hz, ... (8 words)
        detected = f[:, 0] != 0
        start = torch.where(detected, f[:, 1], torch.full_like(f[:, 1], -1)).contiguous()
        cfo = res.view(torch.float32)[:, 3]
        if known_cfo_hz is not None:                    # last_cfo = known + residual when |known| > 0.1 (:275-281)
            cfo = torch.where(known_cfo_hz.abs() > 0.1, known_cfo_hz + cfo, cfo)
        cfo = torch.where(detected, cfo, torch.zeros_like(cfo)).contiguous()
        n_llr = self.config.soft_bits_per_frame(frame_len)
        stride = max(4, (n_llr + 3) & ~3)
        llr = torch.empty((n, stride), dtype=torch.float32, device=dev)
        cnt = torch.empty((n,), dtype=torch.int32, device=dev)
        ctx = self.dem.ctx
        ctx.set_stream(torch.cuda.current_stream(dev))
        ctx.check(lib().ria_mcdpsk_process_batch_at_dev(
            ctx.handle, C.addressof(self.config), _ptr(rows), rows.stride(0), int(frame_len), _ptr(start), _ptr(cfo),
            C.c_void_p(0), n, _ptr(llr), stride, _ptr(cnt), C.c_void_p(0), C.c_void_p(0)))
        slots = n_llr // 648
        return self.decoder.decode_batch(llr[:, : slots * 648]), res

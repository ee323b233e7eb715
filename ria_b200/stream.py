"""One step of the reference's receive state machine for a batch of receptions:
StreamingDecoder::decodeCurrentFrame (src/gui/modem/streaming_decoder.cpp:1060-2125).

The reference runs this once sync has been found and enough samples have arrived: it copies a control-sized (or, after
an escalation, an exactly sized) frame from its ring buffer, classifies it (PING energy test, control-first peek),
demodulates, peeks at codeword 0, either asks for more samples (`pending_total_cw`) or decodes the frame, and falls back
on a few recovery steps when the decode fails.  Here the same decisions are taken for many receptions at once: every
demodulation and every LDPC decode is a device launch over the receptions that reach the step, the branching between
the steps is host logic like the reference's.  A reception is a window of samples plus the sync position inside it; the
caller repeats the step with the returned `pending_total_cw` when the state says SYNC_FOUND, exactly as
checkIfReadyToDecode hands the frame back to decodeCurrentFrame.

Built: the connected OFDM_CHIRP branch complete (control-first peek at the DQPSK R1/4 profile, CFO feedback with the 2 Hz
drift clamp, the R1/4 / data-rate codeword-0 peek, QAM partial-frame escalation, decodeFrame, small-frame recovery, the
multi-candidate light-sync recovery) and the PING energy test a disconnected receiver classifies chirp-only
transmissions with (`ping_energy_batch`), and burst-interleaved groups (`OfdmConnectedStep.burst_group`).  Not built:
CSS frame typing and the ring buffer with its clocks (burst timeout); the MC-DPSK branch of the step, its
disconnected-handshake retries included, is `McdpskStep`.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from ._lib import Context, RiaError, lib
from . import fec, ofdm
from .mcdpsk import _CONTROL_TYPES, _crc16_rows

SEARCHING, SYNC_FOUND, DECODING, BURST_ACCUMULATING = 0, 1, 2, 3
LDPC_BLOCK = 648
R1_4 = 0


def _parse_header(d0: np.ndarray):
    """v2::parseHeader on rows of >= 20 codeword-0 bytes -> (valid, type, total_cw)"""
    ftype = d0[:, 2].astype(np.int32)
    is_control = np.isin(ftype, _CONTROL_TYPES)
    crc_ctl = _crc16_rows(d0[:, :18]) == ((d0[:, 18].astype(np.uint16) << 8) | d0[:, 19])
    crc_dat = _crc16_rows(d0[:, :15]) == ((d0[:, 15].astype(np.uint16) << 8) | d0[:, 16])
    valid = (d0[:, 0] == 0x55) & (d0[:, 1] == 0x4C) & np.where(is_control, crc_ctl, crc_dat)
    total_cw = np.where(is_control, 1, d0[:, 12].astype(np.int32))
    return valid, ftype, total_cw, is_control


def _initial_phase(cfo_hz: np.ndarray, ref_sample: np.ndarray, sample_rate: float) -> np.ndarray:
    """OFDMChirpWaveform::process (ofdm_chirp_waveform.cpp:402-411): a double expression rounded to float, then wrapped
    to [-pi, pi] in double steps rounded to float"""
    out = np.zeros(len(cfo_hz), np.float32)
    for i in range(len(cfo_hz)):
        p = np.float32(-2.0 * np.pi * float(cfo_hz[i]) * float(ref_sample[i]) / float(sample_rate))
        while float(p) > np.pi:
            p = np.float32(float(p) - 2.0 * np.pi)
        while float(p) < -np.pi:
            p = np.float32(float(p) + 2.0 * np.pi)
        out[i] = p
    return out


class OfdmConnectedStep:
    """decodeCurrentFrame for connected OFDM_CHIRP receivers that run `modulation` / `rate` as their data profile."""

    def __init__(self, modulation: int, rate: int, ctx: Optional[Context] = None):
        self.ctx = ctx
        self.modulation, self.rate = int(modulation), int(rate)
        self.cfg = ofdm.ModemConfig.for_waveform(modulation, rate)                 # waveform_->configure(mod, rate)
        self.cfg_ctl = ofdm.ModemConfig.for_waveform(ofdm.DQPSK, R1_4)             # the control profile (:1281-1283)
        self.dem = ofdm.OFDMDemodulator(self.cfg, ctx)
        self.dem_ctl = ofdm.OFDMDemodulator(self.cfg_ctl, ctx)
        self.sym = self.cfg.getSymbolDuration()
        self.data_carriers = self.cfg.getDataCarriers()
        self.bits_per_symbol = self.data_carriers * ofdm.getBitsPerSymbol(modulation)
        self.decoder = ofdm.OfdmFrameDecoder(modulation, rate, self.data_carriers, connected=True,
                                             use_channel_interleave=True, ctx=ctx)
        self.robust_r14 = fec.LDPCDecoder(R1_4, ctx)
        self.robust_rate = fec.LDPCDecoder(rate, ctx)
        self.bpc = fec.code_params(rate)[0] // 8

    # ---- sizing (ofdm_chirp_waveform.cpp getMinSamplesForCWCount; streaming_decoder.cpp:42-68) ----
    def samples_for_cw(self, n_cw: int) -> int:
        data_symbols = (n_cw * LDPC_BLOCK + self.bits_per_symbol - 1) // self.bits_per_symbol
        return (2 + data_symbols) * self.sym

    def control_samples(self) -> int:
        default = self.samples_for_cw(1)
        if self.modulation == ofdm.DQPSK and self.rate == R1_4:
            return default
        carriers = int(self.cfg.num_carriers)
        pilot_count = (carriers + 10 - 1) // 10
        bits = max(1, carriers - pilot_count) * 2
        return max(default, (2 + (LDPC_BLOCK + bits - 1) // bits) * self.sym)

    # ---- waveform_->setFrequencyOffset(cfo); waveform_->process(frame) for a subset ----
    def _process(self, dem, window, idx, pos, length, cfo, phase_ref=None, undo_marker=False):
        """-> (ready bool[m], soft (CUDA [m, stride]), n_soft int[m], cfo_out f32[m], fading f32[m]).  phase_ref: the
        sample index the initial CFO phase refers to (default: pos); undo_marker: restore a negated first LTS symbol."""
        m = len(idx)
        dev = window.device
        L = window.shape[1]
        ready = np.zeros(m, bool)
        n_soft = np.zeros(m, np.int32)
        cfo_out = cfo.astype(np.float32).copy()
        fading = np.zeros(m, np.float32)
        soft = None
        lens = np.minimum(length, L - pos)
        for ln in np.unique(lens):
            sel = np.nonzero(lens == ln)[0]
            if ln < self.sym:
                continue                                                       # process() returns false (:463 of the adapter)
            rows = torch.from_numpy(idx[sel]).to(dev)
            cols = torch.from_numpy(pos[sel].astype(np.int64)).to(dev)[:, None] + torch.arange(int(ln), device=dev)[None, :]
            frames = torch.gather(window.index_select(0, rows), 1, cols)
            if undo_marker:                                                    # ofdm_chirp_waveform.cpp:423-436
                frames[:, : self.sym] = -frames[:, : self.sym]
            phase = _initial_phase(cfo[sel], (pos if phase_ref is None else phase_ref)[sel], float(dem.config.sample_rate))
            out = dem.process_presynced_batch(frames, torch.from_numpy(cfo[sel].astype(np.float32)).to(dev),
                                              torch.from_numpy(phase).to(dev))
            ns = out["n_llr"].cpu().numpy()
            if soft is None or soft.shape[1] < out["llr"].shape[1]:
                grown = torch.zeros((m, out["llr"].shape[1]), dtype=torch.float32, device=dev)
                if soft is not None:
                    grown[:, : soft.shape[1]] = soft
                soft = grown
            soft[torch.from_numpy(sel).to(dev), : out["llr"].shape[1]] = out["llr"]
            n_soft[sel] = ns
            ready[sel] = ns >= LDPC_BLOCK
            cfo_out[sel] = out["cfo"].cpu().numpy()
            fading[sel] = out["fading"].cpu().numpy()
        if soft is None:
            soft = torch.zeros((m, 4), dtype=torch.float32, device=dev)
        return ready, soft, n_soft, cfo_out, fading

    def _robust_cw0(self, dec, soft):
        info, ok, _, _ = dec.robust_decode_batch(soft[:, :LDPC_BLOCK].contiguous(), info_stride=64)
        return info.cpu().numpy(), ok.cpu().numpy().astype(bool)

    def _decode_frame(self, soft, n_soft):
        """decodeFrame on receptions whose soft-bit counts may differ: grouped by count"""
        m = soft.shape[0]
        res = None
        for ns in np.unique(n_soft):
            sel = np.nonzero(n_soft == ns)[0]
            r = self.decoder.decode_batch(soft.index_select(0, torch.from_numpy(sel).to(soft.device))[:, : int(ns)].contiguous())
            if res is None:
                res = {k: (np.zeros((m,) + v.shape[1:], v.dtype) if k != "frame" else np.zeros((m, 1024), np.uint8)) for k, v in r.items()}
            for k, v in r.items():
                if k == "frame":
                    res[k][sel, : v.shape[1]] = v
                else:
                    res[k][sel] = v
        return res

    def step(self, window: torch.Tensor, sync_pos, sync_cfo, last_cfo, pending_total_cw=None):
        """window: CUDA fp32 [n, L]; sync_pos int[n] (first LTS sample); sync_cfo / last_cfo f32[n] (sync_cfo_ / last_cfo_
        of each receiver); pending_total_cw int[n] (0 = first pass).  Returns numpy arrays: state, pending_total_cw,
        has_frame, success, frame_type, codewords_ok, codewords_failed, frame_len, frame u8[n, 1024], last_cfo, sync_pos,
        consumed_len."""
        if not (isinstance(window, torch.Tensor) and window.is_cuda and window.dtype == torch.float32 and window.dim() == 2):
            raise RiaError("step wants CUDA fp32 [n, L] windows (no CPU fallback)")
        n, L = window.shape
        sync_pos = np.asarray(sync_pos, np.int64).copy()
        sync_cfo = np.asarray(sync_cfo, np.float32).copy()
        last_cfo = np.asarray(last_cfo, np.float32).copy()
        pending = np.zeros(n, np.int32) if pending_total_cw is None else np.asarray(pending_total_cw, np.int32).copy()
        out = dict(state=np.full(n, SEARCHING, np.int32), pending_total_cw=pending.copy(), has_frame=np.zeros(n, np.uint8),
                   success=np.zeros(n, np.uint8), frame_type=np.zeros(n, np.int32), codewords_ok=np.zeros(n, np.int32),
                   codewords_failed=np.zeros(n, np.int32), frame_len=np.zeros(n, np.int32),
                   frame=np.zeros((n, 1024), np.uint8), last_cfo=last_cfo, sync_pos=sync_pos,
                   consumed_len=np.zeros(n, np.int64))
        ctl_len = self.control_samples()
        frame_len = np.where(pending > 0, [self.samples_for_cw(int(p)) if p > 0 else 0 for p in pending], ctl_len).astype(np.int64)
        frame_len = np.minimum(frame_len, L - sync_pos)
        out["consumed_len"][:] = frame_len
        alive = frame_len > 0                                                  # empty frame buffer -> SEARCHING (:1118-1125)
        all_idx = np.arange(n)

        def finish(idx, res_rows, res):
            """queue the DecodeResult rows `res_rows` of `res` for receptions idx"""
            for k in ("success", "frame_type", "codewords_ok", "codewords_failed", "frame_len"):
                out[k][idx] = res[k][res_rows]
            out["frame"][idx] = res["frame"][res_rows]
            out["has_frame"][idx] = (res["success"][res_rows] != 0) | (res["codewords_ok"][res_rows] > 0)

        # ---- control-first hypothesis (:1271-1344) ----
        peek = np.nonzero(alive & (pending == 0) & (frame_len <= ctl_len))[0]
        if len(peek):
            ready, soft, n_soft, _, _ = self._process(self.dem_ctl, window, peek, sync_pos[peek], frame_len[peek], sync_cfo[peek])
            cand = np.nonzero(ready)[0]
            if len(cand):
                info, ok = self._robust_cw0(self.robust_r14, soft.index_select(0, torch.from_numpy(cand).to(window.device)))
                d = info[:, :20].copy()
                valid, ftype, total_cw, is_ctl = _parse_header(d)
                hit = np.nonzero(ok & valid & (total_cw == 1) & is_ctl)[0]
                g = peek[cand[hit]]
                out["success"][g] = 1
                out["has_frame"][g] = 1
                out["frame_type"][g] = ftype[hit]
                out["codewords_ok"][g] = 1
                out["frame_len"][g] = 20
                out["frame"][g, :20] = d[hit]
                alive[g] = False
        # ---- data profile (:1346-1378) ----
        idx = np.nonzero(alive)[0]
        if len(idx) == 0:
            return out
        ready, soft, n_soft, cfo_est, _ = self._process(self.dem, window, idx, sync_pos[idx], frame_len[idx], sync_cfo[idx])
        # process() failed / no soft bits -> back to SEARCHING without a frame
        keep = np.nonzero(ready)[0]
        idx, soft, n_soft, cfo_est = idx[keep], soft.index_select(0, torch.from_numpy(keep).to(window.device)), n_soft[keep], cfo_est[keep]
        if len(idx) == 0:
            return out
        # ---- pilot-corrected CFO feedback, drift clamped to 2 Hz while connected (:1411-1434) ----
        cur = last_cfo[idx]
        drift = (cfo_est - cur).astype(np.float32)
        clamp = np.abs(drift) > np.float32(2.0)
        corrected = np.where(clamp, (cur + np.copysign(np.float32(2.0), drift)).astype(np.float32), cfo_est).astype(np.float32)
        last_cfo[idx] = corrected
        sync_cfo[idx] = corrected
        # ---- codeword-0 peek on a one-codeword buffer (:1505-1572) ----
        go = np.ones(len(idx), bool)                                            # receptions that reach decodeFrame
        one = np.nonzero((pending[idx] == 0) & (n_soft >= LDPC_BLOCK) & (n_soft < 2 * LDPC_BLOCK))[0]
        if len(one):
            sub = soft.index_select(0, torch.from_numpy(one).to(window.device))
            info, ok = self._robust_cw0(self.robust_r14, sub)
            d = info[:, :20].copy()
            valid, _, total_cw, _ = _parse_header(d)
            magic = ok & (d[:, 0] == 0x55) & (d[:, 1] == 0x4C)
            fell = magic & valid & (total_cw == 1)
            esc = magic & valid & (total_cw > 1)
            new_pending = np.where(esc, total_cw, 0).astype(np.int32)
            unresolved = ~fell & ~esc
            if self.rate != R1_4 and unresolved.any():
                u = np.nonzero(unresolved)[0]
                info2, ok2 = self._robust_cw0(self.robust_rate, sub.index_select(0, torch.from_numpy(u).to(window.device)))
                d2 = info2[:, : max(self.bpc, 20)].copy()
                d2[:, self.bpc:] = 0
                v2_, _, t2, _ = _parse_header(d2)
                m2 = ok2 & (d2[:, 0] == 0x55) & (d2[:, 1] == 0x4C)
                fell[u] = m2 & v2_ & (t2 == 1)
                new_pending[u] = np.where(m2 & v2_ & (t2 > 1), t2, np.where(m2 & ~(v2_ & (t2 == 1)), 4, 0))
                # decode failed at both rates -> four codewords (:1566-1571)
                new_pending[u] = np.where(~m2, 4, new_pending[u])
            else:
                new_pending = np.where(unresolved, 4, new_pending).astype(np.int32)
            escal = ~fell
            g = idx[one[escal]]
            out["state"][g] = SYNC_FOUND
            out["pending_total_cw"][g] = new_pending[escal]
            go[one[escal]] = False
        # ---- QAM partial frame (:1580-1596) ----
        if self.modulation in (ofdm.QAM16, ofdm.QAM32, ofdm.QAM64, ofdm.QAM256):
            part = go & (pending[idx] == 0) & (n_soft >= 2 * LDPC_BLOCK) & (n_soft < 4 * LDPC_BLOCK)
            g = idx[part]
            out["state"][g] = SYNC_FOUND
            out["pending_total_cw"][g] = 4
            go &= ~part
        sel = np.nonzero(go)[0]
        if len(sel) == 0:
            return out
        idx, n_soft = idx[sel], n_soft[sel]
        soft = soft.index_select(0, torch.from_numpy(sel).to(window.device))
        res = self._decode_frame(soft, n_soft)
        finish(idx, np.arange(len(idx)), res)
        # ---- small-frame recovery (:1803-1852) ----
        failed = np.nonzero(res["success"] == 0)[0]
        one_cw = self.samples_for_cw(1)
        failed = failed[frame_len[idx[failed]] >= one_cw]
        if len(failed):
            g = idx[failed]
            rdy, sbits, ns, _, _ = self._process(self.dem, window, g, sync_pos[g], np.full(len(g), one_cw, np.int64), sync_cfo[g])
            todo = np.nonzero(ns >= LDPC_BLOCK)[0]
            for dec, bpc in ((self.robust_r14, 20), (self.robust_rate, self.bpc)):
                if len(todo) == 0 or (dec is self.robust_rate and self.rate == R1_4):
                    break
                sub = sbits.index_select(0, torch.from_numpy(todo).to(window.device))
                info, ok = self._robust_cw0(dec, sub)
                d = info[:, : max(bpc, 20)].copy()
                d[:, bpc:] = 0
                valid, _, total_cw, _ = _parse_header(d)
                magic = ok & (d[:, 0] == 0x55) & (d[:, 1] == 0x4C)
                single = np.nonzero(magic & valid & (total_cw == 1))[0]
                if len(single):
                    r1 = self._decode_frame(sub.index_select(0, torch.from_numpy(single).to(window.device)), ns[todo[single]])
                    finish(g[todo[single]], np.arange(len(single)), r1)
                    out["consumed_len"][g[todo[single]]] = one_cw
                multi = np.nonzero(magic & valid & (total_cw > 1) & (total_cw < 4))[0]
                for j in multi:                                                # 2-3 codeword frames: reprocess at the exact size
                    gi = g[todo[j]: todo[j] + 1]
                    exact = min(self.samples_for_cw(int(total_cw[j])), int(frame_len[gi[0]]))
                    _, sb, nsb, _, _ = self._process(self.dem, window, gi, sync_pos[gi], np.array([exact], np.int64), sync_cfo[gi])
                    r2 = self._decode_frame(sb, nsb)
                    finish(gi, np.arange(1), r2)
                    out["consumed_len"][gi] = exact
                handled = np.zeros(len(todo), bool)
                handled[single] = True
                handled[multi] = True
                todo = todo[~handled]
        # ---- multi-candidate light-sync recovery (:1859-1962) ----
        retry = idx[(out["success"][idx] == 0) & (out["codewords_ok"][idx] == 0)]
        for delta in (8, -8, 16, -16, 24, -24, 32, -32):
            if len(retry) == 0:
                break
            pos = sync_pos[retry] + delta
            ok_pos = (pos >= 0) & (pos < L)
            cand = retry[ok_pos]
            if len(cand) == 0:
                continue
            ln = np.minimum(out["consumed_len"][cand], L - pos[ok_pos])
            rdy, sb, nsb, cfo_r, _ = self._process(self.dem, window, cand, pos[ok_pos], ln, sync_cfo[cand])
            use = np.nonzero(rdy & (nsb > 0))[0]
            if len(use) == 0:
                continue
            r3 = self._decode_frame(sb.index_select(0, torch.from_numpy(use).to(window.device)), nsb[use])
            good = np.nonzero((r3["success"] != 0) | (r3["codewords_ok"] > 0))[0]
            if len(good) == 0:
                continue
            g = cand[use[good]]
            finish(g, good, r3)
            cur = last_cfo[g]
            est = cfo_r[use[good]]
            drift = (est - cur).astype(np.float32)
            clamp = np.abs(drift) > np.float32(2.0)
            corrected = np.where(clamp, (cur + np.copysign(np.float32(2.0), drift)).astype(np.float32), est).astype(np.float32)
            last_cfo[g] = corrected
            sync_pos[g] = pos[ok_pos][use[good]]
            out["consumed_len"][g] = ln[use[good]]
            retry = np.setdiff1d(retry, g)
        return out

    def burst_group(self, window: torch.Tensor, sync_pos, sync_cfo, last_cfo, group_size: int = 4):
        """A burst-interleaved group (decodeCurrentFrame with the burst marker latched, :1380-1409, then
        accumulateBurstFrames / tryDemodulateNextBurstFrame / finalizeBurstGroup, :3065-3240) for receptions whose
        light sync reported the marker (negated first LTS): `group_size` frames back to back from sync_pos, the first
        demodulated with the marker undone, every further block after an energy check, the CFO fed back from frame to
        frame with the 2 Hz clamp, then BurstInterleaver::deinterleave and decodeFrame of every logical frame.
        Returns (results, last_cfo): results[k] = dict of numpy arrays for logical frame k (success, frame_type,
        codewords_ok, codewords_failed, frame_len, frame, queued) with queued = the reference pushes the result
        (success or a codeword decoded); a group that was aborted has queued = 0 everywhere."""
        n, L = window.shape
        dev = window.device
        sync_pos = np.asarray(sync_pos, np.int64)
        burst_cfo = np.asarray(sync_cfo, np.float32).copy()
        last_cfo = np.asarray(last_cfo, np.float32).copy()
        G = max(2, int(group_size))
        block = self.samples_for_cw(4)                                          # burst_min_block_ = getMinSamplesForFrame()
        alive = np.ones(n, bool)
        soft_all = torch.zeros((n, G, 2592), dtype=torch.float32, device=dev)
        all_idx = np.arange(n)

        def clamp_to(cur, est):
            drift = (est - cur).astype(np.float32)
            big = np.abs(drift) > np.float32(2.0)
            return np.where(big, (cur + np.copysign(np.float32(2.0), drift)).astype(np.float32), est).astype(np.float32)

        for k in range(G):
            idx = all_idx[alive]
            if len(idx) == 0:
                break
            pos = sync_pos[idx] + k * block
            enough = pos + block <= L
            if k > 0:
                # a block that has not arrived yet: the reference waits, then times out and discards the group
                alive[idx[~enough]] = False
                idx, pos = idx[enough], pos[enough]
                if len(idx) == 0:
                    break
                # energy check on [1024, 1024 + 5000) of the block (:3153-3169)
                cols = torch.from_numpy(pos).to(dev)[:, None] + torch.arange(1024, min(block, 1024 + 5000), device=dev)[None, :]
                seg = torch.gather(window.index_select(0, torch.from_numpy(idx).to(dev)), 1, cols).contiguous()
                e = ping_energy_batch(seg, 0, self.ctx)                         # training_skip 0: the data RMS is the block's
                lost = e["data_rms"] < np.float32(0.04)
                alive[idx[lost]] = False
                idx, pos = idx[~lost], pos[~lost]
                if len(idx) == 0:
                    break
            ready, soft, n_soft, cfo_est, _ = self._process(self.dem, window, idx, pos, np.full(len(idx), block, np.int64),
                                                            burst_cfo[idx], phase_ref=sync_pos[idx] if k else None,
                                                            undo_marker=(k == 0))
            ok = ready & (n_soft >= 2592)
            alive[idx[~ok]] = False
            keep = np.nonzero(ok)[0]
            idx = idx[keep]
            if len(idx) == 0:
                break
            soft_all[torch.from_numpy(idx).to(dev), k] = soft.index_select(0, torch.from_numpy(keep).to(dev))[:, :2592]
            cur = last_cfo[idx] if k == 0 else burst_cfo[idx]                   # frame 0 clamps against last_cfo_ (:1394-1402)
            corrected = clamp_to(cur, cfo_est[keep])
            burst_cfo[idx] = corrected
            last_cfo[idx] = corrected
        results = []
        done = all_idx[alive]
        logical = fec.burst_deinterleave_batch(soft_all.index_select(0, torch.from_numpy(done).to(dev)), self.ctx) if len(done) else None
        for k in range(G):
            r = dict(success=np.zeros(n, np.uint8), frame_type=np.zeros(n, np.int32), codewords_ok=np.zeros(n, np.int32),
                     codewords_failed=np.zeros(n, np.int32), frame_len=np.zeros(n, np.int32), frame=np.zeros((n, 1024), np.uint8),
                     queued=np.zeros(n, np.uint8))
            if len(done):
                d = self.decoder.decode_batch(logical[:, k].contiguous())
                for key in ("success", "frame_type", "codewords_ok", "codewords_failed", "frame_len"):
                    r[key][done] = d[key]
                r["frame"][done, : d["frame"].shape[1]] = d["frame"]
                r["queued"][done] = (d["success"] != 0) | (d["codewords_ok"] > 0)
            results.append(r)
        return results, last_cfo


class McdpskStep:
    """decodeCurrentFrame for MC-DPSK receivers (streaming_decoder.cpp:1060-1644): one-codeword frame buffer, PING energy
    test (disconnected receivers), process, the codeword-0 peek with the handshake rules (a disconnected receiver waits
    for at least the three codewords of a CONNECT before it decodes), the early-window rule, decodeFrame
    (= decodeMCDPSKFrame), the header salvage, and the retries a disconnected receiver makes when the handshake decode
    fails outright: once with the other differential modulation (:1646-1690), then at the neighbouring sync offsets
    +-8 .. +-64 with both modulations (:1692-1795)."""

    PING, CONNECT_PAYLOAD = 0x01, 25                       # v2::FrameType::PING; ConnectFrame::PAYLOAD_SIZE (frame_v2.hpp:556)

    def __init__(self, config, connected: bool, rate: int = R1_4, ctx: Optional[Context] = None, chase_cache=None):
        from . import mcdpsk
        self.config, self.connected, self.ctx = config, bool(connected), ctx
        self.rate = int(rate) if connected else R1_4                              # :1437
        self.dem = mcdpsk.MCDPSKDemodulator(config, ctx)
        self.decoder = mcdpsk.McdpskFrameDecoder(self.rate, ctx, chase_cache)
        self.robust = fec.LDPCDecoder(self.rate, ctx)
        self.bpc = fec.code_params(self.rate)[0] // 8
        k = fec.code_params(self.rate)[0]
        # min_handshake_cw = max(2, DataFrame::calculateCodewords(ConnectFrame::PAYLOAD_SIZE, rate)) (:1448-1451)
        self.min_handshake_cw = 0 if connected else max(2, ((17 + self.CONNECT_PAYLOAD + 2) * 8 + k - 1) // k)
        self.min_handshake_cw_r14 = max(2, ((17 + self.CONNECT_PAYLOAD + 2) * 8 + 162 - 1) // 162)
        # the other differential modulation, for the disconnected-handshake retries (DBPSK <-> DQPSK, :1650-1660)
        self.dem_alt = None
        if not connected and int(config.bits_per_symbol) in (1, 2):
            alt = type(config).from_buffer_copy(bytes(config))
            alt.bits_per_symbol = 3 - int(config.bits_per_symbol)
            self.dem_alt = mcdpsk.MCDPSKDemodulator(alt, ctx)

    def samples_for_cw(self, n_cw: int) -> int:               # MCDPSKWaveform::getMinSamplesForCWCount (mc_dpsk_waveform.cpp:470-484)
        c = self.config
        bits = int(c.num_carriers) * int(c.bits_per_symbol)
        per_cw = (LDPC_BLOCK + bits - 1) // bits
        return int(c.training_symbols) * int(c.samples_per_symbol) + int(c.samples_per_symbol) + \
            n_cw * per_cw * int(c.samples_per_symbol) * int(c.spreading)

    def step(self, window: torch.Tensor, sync_pos, sync_cfo, pending_total_cw=None):
        """window CUDA fp32 [n, L]; sync_pos int[n] = training start; sync_cfo f32[n]; pending_total_cw int[n].
        Returns numpy arrays: state, pending_total_cw, has_frame, success, is_ping, frame_type, codewords_ok,
        codewords_failed, frame_len, frame u8[n, W], sync_pos (moved by the handshake sync recovery)."""
        if not (isinstance(window, torch.Tensor) and window.is_cuda and window.dtype == torch.float32 and window.dim() == 2):
            raise RiaError("step wants CUDA fp32 [n, L] windows (no CPU fallback)")
        n, L = window.shape
        dev = window.device
        sync_pos = np.asarray(sync_pos, np.int64)
        sync_cfo = np.asarray(sync_cfo, np.float32)
        pending = np.zeros(n, np.int32) if pending_total_cw is None else np.asarray(pending_total_cw, np.int32).copy()
        out = dict(state=np.full(n, SEARCHING, np.int32), pending_total_cw=pending.copy(), has_frame=np.zeros(n, np.uint8),
                   success=np.zeros(n, np.uint8), is_ping=np.zeros(n, np.uint8), frame_type=np.zeros(n, np.int32),
                   codewords_ok=np.zeros(n, np.int32), codewords_failed=np.zeros(n, np.int32),
                   frame_len=np.zeros(n, np.int32), frame=np.zeros((n, 1024), np.uint8), sync_pos=sync_pos.copy())
        frame_len = np.array([self.samples_for_cw(int(p) if p > 0 else 1) for p in pending], np.int64)
        frame_len = np.minimum(frame_len, L - sync_pos)
        alive = frame_len > 0
        reached = np.zeros(n, bool)                                     # receptions whose decodeFrame ran and settled the step
        for ln in np.unique(frame_len[alive]):
            idx = np.nonzero(alive & (frame_len == ln))[0]
            cols = torch.from_numpy(sync_pos[idx]).to(dev)[:, None] + torch.arange(int(ln), device=dev)[None, :]
            frames = torch.gather(window.index_select(0, torch.from_numpy(idx).to(dev)), 1, cols).contiguous()
            live = np.ones(len(idx), bool)
            # ---- PING energy test, disconnected receivers only (:1127-1266) ----
            if not self.connected:
                e = ping_energy_batch(frames, 4608, self.ctx)
                hit = np.nonzero(e["is_ping"])[0]
                g = idx[hit]
                out["has_frame"][g] = 1; out["success"][g] = 1; out["is_ping"][g] = 1; out["frame_type"][g] = self.PING
                live[hit] = False
            sel = np.nonzero(live)[0]
            if len(sel) == 0:
                continue
            idx2 = idx[sel]
            fr = frames.index_select(0, torch.from_numpy(sel).to(dev))
            # ---- waveform_->setFrequencyOffset(sync_cfo_); process(frame_buffer) ----
            d = self.dem.process_batch(fr, torch.from_numpy(sync_cfo[idx2]).to(dev))
            n_soft = d["n_llr"].cpu().numpy()
            soft = d["llr"]
            okp = n_soft > 0                                                    # process() false / no soft bits -> SEARCHING
            pend = pending[idx2]
            go = okp.copy()
            avail = n_soft // LDPC_BLOCK
            # ---- codeword-0 peek (:1443-1503) ----
            peek = np.nonzero(okp & (pend == 0) & (n_soft >= LDPC_BLOCK))[0]
            if len(peek):
                info, ok, _, _ = self.robust.robust_decode_batch(soft.index_select(0, torch.from_numpy(peek).to(dev))[:, :LDPC_BLOCK].contiguous(),
                                                                 info_stride=64)
                info = info.cpu().numpy(); ok = ok.cpu().numpy().astype(bool)
                dd = info[:, : max(self.bpc, 20)].copy()
                valid, ftype, total_cw, _ = _parse_header(dd)
                magic = ok & (dd[:, 0] == 0x55) & (dd[:, 1] == 0x4C)
                needed = total_cw.copy()
                handshake = np.isin(ftype, (0x12, 0x13, 0x14))                   # isConnectFrame minus DISCONNECT
                if self.min_handshake_cw > 0:
                    needed = np.where(handshake & (needed < self.min_handshake_cw), self.min_handshake_cw, needed)
                esc = magic & valid & (needed > 1) & (avail[peek] < needed)
                new_p = np.where(esc, needed, 0)
                if self.min_handshake_cw > 1:
                    fb = ~esc & (avail[peek] < self.min_handshake_cw)
                    new_p = np.where(fb, self.min_handshake_cw, new_p)
                    esc = esc | fb
                g = idx2[peek[esc]]
                out["state"][g] = SYNC_FOUND
                out["pending_total_cw"][g] = new_p[esc]
                go[peek[esc]] = False
            if not self.connected:
                # early window (:1598-1608) and the full-handshake window (:1610-1624)
                early = go & (pend == 0) & (n_soft < LDPC_BLOCK)
                out["state"][idx2[early]] = SYNC_FOUND
                out["pending_total_cw"][idx2[early]] = 1
                go &= ~early
                hs = go & (avail < self.min_handshake_cw_r14) & (pend < self.min_handshake_cw_r14)
                out["state"][idx2[hs]] = SYNC_FOUND
                out["pending_total_cw"][idx2[hs]] = self.min_handshake_cw_r14
                go &= ~hs
            dec = np.nonzero(go & (n_soft >= LDPC_BLOCK))[0]
            for ns in np.unique(n_soft[dec]):
                sub = dec[n_soft[dec] == ns]
                slots = int(ns) // LDPC_BLOCK
                r = self.decoder.decode_batch(soft.index_select(0, torch.from_numpy(sub).to(dev))[:, : slots * LDPC_BLOCK].contiguous())
                g = idx2[sub]
                for key in ("success", "frame_type", "codewords_ok", "codewords_failed", "frame_len"):
                    out[key][g] = r[key]
                out["frame"][g, : r["frame"].shape[1]] = r["frame"]
                # header salvage (:1631-1644)
                salv = (r["success"] == 0) & (pend[sub] == 0) & (r["codewords_ok"] > 0)
                if salv.any():
                    hv, _, ht, _ = _parse_header(np.pad(r["frame"], ((0, 0), (0, max(0, 20 - r["frame"].shape[1]))))[:, :64])
                    need = salv & hv & (ht > 1) & (avail[sub] < ht)
                    out["state"][g[need]] = SYNC_FOUND
                    out["pending_total_cw"][g[need]] = ht[need]
                    out["codewords_ok"][g[need]] = 0
                queued = (out["state"][g] == SEARCHING) & ((r["success"] != 0) | (r["codewords_ok"] > 0))
                out["has_frame"][g] = queued
                reached[g] = out["state"][g] == SEARCHING
        if not self.connected:
            self._handshake_retries(window, sync_pos, sync_cfo, frame_len, reached, out)
        return out

    def _try(self, dem, window, idx, pos, ln, sync_cfo):
        """process + decodeMCDPSKFrame(R1/4) of receptions idx at positions pos with ln samples -> decode dict or None"""
        dev = window.device
        cols = torch.from_numpy(pos).to(dev)[:, None] + torch.arange(int(ln), device=dev)[None, :]
        frames = torch.gather(window.index_select(0, torch.from_numpy(idx).to(dev)), 1, cols).contiguous()
        d = dem.process_batch(frames, torch.from_numpy(sync_cfo[idx]).to(dev))
        n_soft = d["n_llr"].cpu().numpy()
        ns = int(n_soft[0]) if len(n_soft) else 0                   # same length, same configuration: same count
        if ns < LDPC_BLOCK:
            return None
        return self.decoder.decode_batch(d["llr"][:, : (ns // LDPC_BLOCK) * LDPC_BLOCK].contiguous())

    def _accept(self, out, g, r, rows):
        for key in ("success", "frame_type", "codewords_ok", "codewords_failed", "frame_len"):
            out[key][g] = r[key][rows]
        out["frame"][g] = 0
        out["frame"][g, : r["frame"].shape[1]] = r["frame"][rows]
        out["has_frame"][g] = 1

    def _handshake_retries(self, window, sync_pos, sync_cfo, frame_len, reached, out):
        L = window.shape[1]
        fail = lambda: np.nonzero(reached & (out["success"] == 0) & (out["codewords_ok"] == 0))[0]
        # ---- once more with the other modulation on the same samples (:1646-1690) ----
        if self.dem_alt is not None:
            todo = fail()
            for ln in np.unique(frame_len[todo]):
                idx = todo[frame_len[todo] == ln]
                r = self._try(self.dem_alt, window, idx, sync_pos[idx], ln, sync_cfo)
                if r is not None:
                    good = np.nonzero(r["success"] != 0)[0]
                    self._accept(out, idx[good], r, good)
        # ---- neighbouring sync offsets, both modulations (:1692-1795) ----
        ctl = self.samples_for_cw(1)
        for delta in (8, -8, 16, -16, 24, -24, 32, -32, 48, -48, 64, -64):
            todo = fail()
            if len(todo) == 0:
                break
            pos = sync_pos[todo] + delta
            ln_all = np.minimum(frame_len[todo], L - pos)
            usable = (pos >= 0) & (ln_all >= ctl)
            for dem in (self.dem, self.dem_alt):
                if dem is None:
                    continue
                cur = np.isin(todo, fail()) & usable                    # a hit of the first modulation ends the delta
                for ln in np.unique(ln_all[cur]):
                    sel = np.nonzero(cur & (ln_all == ln))[0]
                    r = self._try(dem, window, todo[sel], pos[sel], ln, sync_cfo)
                    if r is None:
                        continue
                    good = np.nonzero(r["success"] != 0)[0]
                    g = todo[sel[good]]
                    self._accept(out, g, r, good)
                    out["sync_pos"][g] = pos[sel[good]]


def ping_energy_batch(frames: torch.Tensor, training_skip: int = 4608, ctx: Optional[Context] = None) -> np.ndarray:
    """The legacy PING test of decodeCurrentFrame (:1127-1160, 1219-1229) for a batch: frames CUDA fp32 [n, frame_len]
    start at the sync position; RMS of the first `training_skip` samples (4608 for MC-DPSK: training + reference
    symbols; the light-preamble length for OFDM) against the RMS of the <= 5000 samples behind them, both sums in the
    reference's sample order.  Returns a structured array (training_rms, data_rms, ratio, is_ping): a ratio under 0.6
    means the transmission ends with the preamble (PING / PONG)."""
    if not (isinstance(frames, torch.Tensor) and frames.is_cuda and frames.dtype == torch.float32 and frames.dim() == 2):
        raise RiaError("ping_energy_batch wants CUDA fp32 [n, frame_len] (no CPU fallback)")
    frames = frames if frames.stride(1) == 1 else frames.contiguous()
    n, flen = frames.shape
    from .fec import default_context
    ctx = ctx or default_context()
    out = torch.empty((n, 4), dtype=torch.float32, device=frames.device)
    ctx.set_stream(torch.cuda.current_stream(frames.device))
    ctx.check(lib().ria_ping_energy_batch_dev(ctx.handle, C.c_void_p(frames.data_ptr()), frames.stride(0), flen,
                                              int(training_skip), n, C.c_void_p(out.data_ptr())))
    o = out.cpu().numpy()
    res = np.zeros(n, dtype=[("training_rms", np.float32), ("data_rms", np.float32), ("ratio", np.float32), ("is_ping", np.uint8)])
    res["training_rms"], res["data_rms"], res["ratio"] = o[:, 0], o[:, 1], o[:, 2]
    res["is_ping"] = o[:, 3] != 0
    return res

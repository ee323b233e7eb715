"""GPU: wide parity runs -- thousands of frames per BASELINE config and per Watterson preset through
demodulator -> decoder on the device and through the unmodified reference (oracle/_ref) on the SAME received
buffers, reporting the exact mismatch COUNT (which must be zero).

* AWGN: RIA_PARITY_FRAMES (default 4096) frames for C4 (OFDM QAM64 R3/4), C1 (OFDM DQPSK R1/2) and C3
  (MC-DPSK DBPSK 4x at -8 dB, chirp-acquired).
* Fading: RIA_PARITY_FADED (default 1024) frames per preset good / moderate / poor
  (sim::WattersonChannel presets, src/sim/hf_channel.hpp:411-488), faded by the REFERENCE channel so that
  both sides see identical samples; this is where the fade-erasure, fading-index, decision-directed and
  magnitude-interpolation branches of channel_equalizer.cpp:645-1451 switch.

The reference side is fanned out over the host cores (fork; the workers never touch CUDA).  The counts are
also written to gpurun_out/parity_wide.json when that directory exists (copied to profiles/ per round).
"""
import json
import multiprocessing as mp
import os
import time

import numpy as np
import pytest

from oracle.bindings import (BITS_PER_CARRIER, BYTES_PER_CW, D8PSK, DQPSK, QAM64, R1_2, R1_4, R3_4, McdpskConfig,
                             ModemConfig, Ref, WattersonConfig as RefWatt)

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_AWGN = int(os.environ.get("RIA_PARITY_FRAMES", "4096"))
N_FADED = int(os.environ.get("RIA_PARITY_FADED", "1024"))
_REPORT = {}
_G = {}                      # inputs shared with the forked workers (copy on write)


def _cores():
    return max(1, min(32, len(os.sched_getaffinity(0))))


def _report(name, entry):
    _REPORT[name] = entry
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "parity_wide.json"), "w") as f:
            json.dump(_REPORT, f, indent=1, sort_keys=True)
    print(f"parity[{name}]: {json.dumps(entry)}")


def _fan_out(worker, n):
    cores = _cores()
    bounds = np.linspace(0, n, cores + 1).astype(int)
    jobs = [(int(bounds[i]), int(bounds[i + 1])) for i in range(cores) if bounds[i + 1] > bounds[i]]
    with mp.get_context("fork").Pool(len(jobs)) as pool:
        parts = pool.map(worker, jobs, chunksize=1)
    return [r for p in parts for r in p]


# ---------------------------------------------------------------------------------------------
# OFDM
# ---------------------------------------------------------------------------------------------
def _ofdm_ref_worker(span):
    lo, hi = span
    ref = Ref()
    cfg, rate, bps = _G["cfg"], _G["rate"], _G["bps"]
    fade = _G.get("fade")
    out = []
    for i in range(lo, hi):
        rx = _G["rx"][i]
        if fade is not None:                                    # the reference's own channel makes the buffer
            rx = ref.watterson_process(fade, 1000 + i, _G["tx"][i])
        r = ref.ofdm_process_presynced(cfg, rx, float(_G["cfo"][i]), float(_G["phase"][i]))
        if i < _G["full_n"]:
            data, ok = ref.decode_fixed_frame_full(r["soft"], rate, True, bps)
        else:
            data, ok, _ = ref.frame_decode_first_pass(r["soft"], rate, True, bps)
        out.append((rx if fade is not None else None, r["soft"], np.float32(r["snr_db"]), np.float32(r["cfo"]),
                    np.float32(r["fading"]), data, ok))
    return out


def _ofdm_tx(ctx, cfg, rate, n, seed):
    """n distinct frames: payload -> v2 frame -> encodeFixedFrame -> OFDM TX, all on the device (byte / sample
    identical to the reference transmitter, tests/test_tx_gpu.py)."""
    import torch
    from ria_b200 import ofdm, txsynth
    bpc = BYTES_PER_CW[rate]
    rng = np.random.default_rng(seed)
    frames = txsynth.make_data_frames("K1ABC", "W2XYZ", 7, rng.integers(0, 256, size=(n, 4 * bpc - 19 - 2), dtype=np.uint8), bpc)
    rcfg = ofdm.ModemConfig.from_buffer_copy(bytes(cfg))
    bps = cfg.data_carriers() * BITS_PER_CARRIER[cfg.modulation]
    coded = ofdm.encode_fixed_frame_batch(torch.from_numpy(frames).cuda(), rate, True, bps, ctx)
    tx = ofdm.ofdm_tx_frames(rcfg, coded, ctx)
    torch.cuda.synchronize()
    return tx.cpu().numpy(), frames, bps


def _ofdm_case(ctx, name, mod, spacing, rate, snr_db, n, fade_cond=None, cfo_span=0.0):
    import torch
    from ria_b200 import ofdm, sim
    cfg = ModemConfig.make(mod, spacing, 1)
    tx, sent, bps = _ofdm_tx(ctx, cfg, rate, n, seed=sum(map(ord, name)))
    rng = np.random.default_rng(1 + sum(map(ord, name)))
    cfo = rng.uniform(-cfo_span, cfo_span, size=n).astype(np.float32) if cfo_span else np.zeros(n, np.float32)
    phase = rng.uniform(-3.1, 3.1, size=n).astype(np.float32) if cfo_span else np.zeros(n, np.float32)
    _G.clear()
    # The reference's retry ladder costs ~0.3 s per failing frame on a host core: on the fading channels,
    # where most QAM64 frames fail, the complete decodeFixedFrame is compared on the first 256 frames and
    # the first pass on all of them; on AWGN every frame goes through the complete decode.
    full_n = n if fade_cond is None else min(n, int(os.environ.get("RIA_PARITY_FADED_FULL", "256")))
    _G.update(cfg=cfg, rate=rate, bps=bps, cfo=cfo, phase=phase, tx=tx, full_n=full_n)
    if fade_cond is None:
        p = np.mean(tx.astype(np.float64) ** 2, axis=1, keepdims=True)
        sigma = np.sqrt(p / 10 ** (snr_db / 10)).astype(np.float32)
        rx = (tx + rng.standard_normal(tx.shape, dtype=np.float32) * sigma).astype(np.float32)
        _G["rx"] = rx
    else:
        _G["rx"] = [None] * n
        _G["fade"] = RefWatt.from_buffer_copy(bytes(sim.WattersonConfig.preset(fade_cond, snr_db)))
    t0 = time.perf_counter()
    want = _fan_out(_ofdm_ref_worker, n)
    t_ref = time.perf_counter() - t0
    if fade_cond is not None:
        rx = np.stack([w[0] for w in want])
    # device: demodulator, then the complete decodeFixedFrame (first pass + retry ladder + repair)
    dem = ofdm.OFDMDemodulator(ofdm.ModemConfig.from_buffer_copy(bytes(cfg)), ctx)
    out = dem.process_presynced_batch(torch.from_numpy(rx).cuda(), torch.from_numpy(cfo).cuda(), torch.from_numpy(phase).cuda())
    data, status = ofdm.decode_fixed_frame_batch(out["llr"][:full_n], rate, True, bps, ctx, retry_ladder=True, fp_repair=True)
    if full_n < n:
        data1, status1 = ofdm.decode_fixed_frame_batch(out["llr"][full_n:], rate, True, bps, ctx, retry_ladder=False, fp_repair=False)
        data, status = torch.cat([data, data1]), torch.cat([status, status1])
    torch.cuda.synchronize()
    llr, n_llr = out["llr"].cpu().numpy(), out["n_llr"].cpu().numpy()
    snr, cfo_o, fad = out["snr_db"].cpu().numpy(), out["cfo"].cpu().numpy(), out["fading"].cpu().numpy()
    data = data.cpu().numpy()
    st = ofdm.status_array(status)
    bad = dict(soft_frames=0, soft_bits=0, n_llr=0, snr=0, cfo=0, fading=0, cw_ok=0, data=0)
    ulp = dict(snr_db_last_bit=0)
    decoded = payload_ok = 0
    for i, (_, w_soft, w_snr, w_cfo, w_fad, w_data, w_ok) in enumerate(want):
        if int(n_llr[i]) != len(w_soft):
            bad["n_llr"] += 1
            continue
        diff = llr[i, : len(w_soft)].view(np.uint32) != w_soft.view(np.uint32)
        bad["soft_frames"] += int(diff.any())
        bad["soft_bits"] += int(diff.sum())
        # the SNR report is 10 log10 of a bit-identical linear estimate; CUDA's log10f and glibc's differ in the last
        # bit on ~10 % of arguments, so the dB figure is held to 1e-6 relative and the last-bit count reported
        bad["snr"] += int(abs(float(snr[i]) - float(w_snr)) > 1e-6 * max(1.0, abs(float(w_snr))))
        ulp["snr_db_last_bit"] += int(np.float32(snr[i]).view(np.uint32) != w_snr.view(np.uint32))
        bad["cfo"] += int(np.float32(cfo_o[i]).view(np.uint32) != w_cfo.view(np.uint32))
        bad["fading"] += int(np.float32(fad[i]).view(np.uint32) != w_fad.view(np.uint32))
        bad["cw_ok"] += int(not np.array_equal(st["cw_ok"][i], w_ok))
        bad["data"] += int(not np.array_equal(data[i], w_data))
        decoded += int(w_ok.all())
        payload_ok += int(w_ok.all() and np.array_equal(w_data[: sent.shape[1]], sent[i]))
    entry = dict(frames=n, frames_full_decode=full_n, mismatches=bad, last_bit_differences=ulp, ref_all_cw_decoded=decoded, ref_payload_correct=payload_ok,
                 ladder_frames=int((st["ladder_cw_mask"] != 0).sum()), repaired_frames=int((st["fp_repair"] == 1).sum()),
                 mean_fading_index=float(np.mean(fad)), frames_fading_above_0p30=int((fad > 0.30).sum()), ref_seconds=round(t_ref, 1), ref_cores=_cores())
    _report(name, entry)
    return entry


def _assert_exact(entry, scalars_exact=True):
    m = entry["mismatches"]
    assert m["n_llr"] == 0 and m["soft_frames"] == 0 and m["soft_bits"] == 0, entry
    assert m["cw_ok"] == 0 and m["data"] == 0, entry
    if scalars_exact:
        assert m["snr"] == 0 and m["cfo"] == 0 and m["fading"] == 0, entry


def test_c4_qam64_awgn_wide(ctx, ref):
    e = _ofdm_case(ctx, "c4_qam64_r34_awgn28", QAM64, 4, R3_4, 28.0, N_AWGN)
    _assert_exact(e)
    assert e["ref_payload_correct"] >= 0.99 * e["frames"]


def test_c4_qam64_awgn_with_sync_cfo_wide(ctx, ref):
    """what production hands over: every frame with the CFO / mixer phase of its sync (cfo ~ U(-5, 5) Hz)"""
    e = _ofdm_case(ctx, "c4_qam64_r34_awgn28_cfo5", QAM64, 4, R3_4, 28.0, max(256, N_AWGN // 4), cfo_span=5.0)
    _assert_exact(e)


def test_c1_dqpsk_awgn_wide(ctx, ref):
    e = _ofdm_case(ctx, "c1_dqpsk_r12_awgn15", DQPSK, 10, R1_2, 15.0, N_AWGN)
    _assert_exact(e)
    assert e["ref_payload_correct"] >= 0.99 * e["frames"]


@pytest.mark.parametrize("cond,cname", [(1, "good"), (2, "moderate"), (3, "poor")])
def test_ofdm_watterson_presets_wide(ctx, ref, cond, cname):
    """frames faded by the reference's WattersonChannel; C4-type coherent QAM64 and C1-type differential DQPSK"""
    e = _ofdm_case(ctx, f"c4_qam64_r34_watterson_{cname}_30dB", QAM64, 4, R3_4, 30.0, N_FADED, fade_cond=cond)
    _assert_exact(e)
    e = _ofdm_case(ctx, f"c1_dqpsk_r12_watterson_{cname}_18dB", DQPSK, 10, R1_2, 18.0, N_FADED, fade_cond=cond)
    _assert_exact(e)


def test_d8psk_single_and_two_pass_wide(ctx, ref):
    """D8PSK (a10): the single-pass demapper on AWGN, and the two-pass demapper that the LTS / pilot fading index
    switches on above 0.30 (demodulator.cpp:286-296, 533-624) on frames faded by the reference channel."""
    e = _ofdm_case(ctx, "d8psk_r12_awgn20", D8PSK, 8, R1_2, 20.0, max(256, N_AWGN // 4))
    _assert_exact(e)
    assert e["mean_fading_index"] < 0.30
    e = _ofdm_case(ctx, "d8psk_r12_watterson_moderate_24dB", D8PSK, 8, R1_2, 24.0, max(256, N_FADED // 2), fade_cond=2)
    _assert_exact(e)
    # the two-pass branch is gated per SYMBOL by last_fading_index > 0.30; frames whose final channel estimate is that
    # selective certainly went through it
    assert e["frames_fading_above_0p30"] >= 20, "operating point does not reach the two-pass branch"


# ---------------------------------------------------------------------------------------------
# MC-DPSK (C3): chirp acquisition -> demodulation at the detected start with the detected CFO -> LDPC
# ---------------------------------------------------------------------------------------------
def _mcdpsk_ref_worker(span):
    lo, hi = span
    ref = Ref()
    cfg, frame_len, window = _G["cfg"], _G["frame_len"], _G["window"]
    fade = _G.get("fade")
    out = []
    for i in range(lo, hi):
        row = _G["rows"][i]
        if fade is not None:
            row = ref.watterson_process(fade, 5000 + i, _G["tx"][i])
        s = ref.chirp_detect_dual(row[:window], 0.15)
        soft, ok, info, it = np.zeros(0, np.float32), 0, np.zeros(24, np.uint8), 0
        start = -1
        if s.detected:
            start = int(s.aux) + 28800                          # training starts one chirp + gap after the down chirp
            if start + frame_len <= len(row):
                r = ref.mcdpsk_process(cfg, row[start:start + frame_len], float(s.cfo_hz))
                soft = r["soft"]
                if len(soft) >= 648:
                    inf, okk, itt = ref.ldpc_decode_batch(R1_4, soft[:648], 50, 0.9375, 24)
                    info, ok, it = inf[0], int(okk[0]), int(itt[0])
        out.append((row if fade is not None else None, int(s.detected), int(s.start_sample), int(s.aux), np.float32(s.cfo_hz),
                    np.float32(s.correlation), soft, ok, info, it))
    return out


def _mcdpsk_case(ctx, name, snr_db, n, fade_cond=None):
    import torch
    from ria_b200 import mcdpsk, sim, sync, txsynth
    cfg = McdpskConfig.make(1, 4, 10)
    rcfg = mcdpsk.MultiCarrierDPSKConfig.from_buffer_copy(bytes(cfg))
    rng = np.random.default_rng(sum(map(ord, name)))
    info_bits = rng.integers(0, 2, size=(n, 162), dtype=np.uint8)
    cw = np.packbits(txsynth.ldpc_encode_bits(info_bits, R1_4), axis=1)
    body = mcdpsk.mcdpsk_tx_frames(rcfg, torch.from_numpy(cw).cuda(), ctx)
    torch.cuda.synchronize()
    body = body.cpu().numpy()
    frame_len = body.shape[1]
    pre = sync.chirp_generate_host()
    lead = rng.integers(500, 4000, size=n)
    window = 120000
    row_len = 4000 + len(pre) + frame_len + 800
    tx = np.zeros((n, row_len), np.float32)
    for i in range(n):
        tx[i, lead[i]: lead[i] + len(pre)] = pre
        tx[i, lead[i] + len(pre): lead[i] + len(pre) + frame_len] = body[i]
    _G.clear()
    _G.update(cfg=cfg, frame_len=frame_len, window=window, tx=tx)
    if fade_cond is None:
        p = np.mean(tx.astype(np.float64) ** 2, axis=1, keepdims=True)
        sigma = np.sqrt(p / 10 ** (snr_db / 10)).astype(np.float32)
        rows = (tx + rng.standard_normal(tx.shape, dtype=np.float32) * sigma).astype(np.float32)
        _G["rows"] = rows
    else:
        _G["rows"] = [None] * n
        _G["fade"] = RefWatt.from_buffer_copy(bytes(sim.WattersonConfig.preset(fade_cond, snr_db)))
    t0 = time.perf_counter()
    want = _fan_out(_mcdpsk_ref_worker, n)
    t_ref = time.perf_counter() - t0
    if fade_cond is not None:
        rows = np.stack([w[0] for w in want])
    chain = mcdpsk.McdpskRxChain(rcfg, R1_4, 50, 0.9375, 0.15, ctx)
    bad = dict(detected=0, start=0, cfo=0, corr_1e4=0, soft_frames=0, soft_bits=0, ok=0, info=0, iters=0)
    decoded = 0
    step = 1024
    for off in range(0, n, step):
        m = min(step, n - off)
        out = chain.process_batch(torch.from_numpy(rows[off:off + m]).cuda(), frame_len, window)
        torch.cuda.synchronize()
        sy = np.frombuffer(out["sync"].cpu().numpy().tobytes(), dtype=sync.SYNC_RESULT_DTYPE)
        llr = out["acc"].cpu().numpy()          # first reception: the accumulator IS the frame's soft bits
        okd, info, iters = out["ok"].cpu().numpy(), out["info"].cpu().numpy(), out["iters"].cpu().numpy()
        for j in range(m):
            _, w_det, w_start, w_aux, w_cfo, w_corr, w_soft, w_ok, w_info, w_it = want[off + j]
            if int(sy["detected"][j]) != w_det:
                bad["detected"] += 1
                continue
            bad["corr_1e4"] += int(abs(float(sy["correlation"][j]) - float(w_corr)) > 1e-4 * max(abs(float(w_corr)), 1e-3))
            if not w_det:
                continue
            bad["start"] += int(sy["start_sample"][j] != w_start or sy["aux"][j] != w_aux)
            bad["cfo"] += int(np.float32(sy["cfo_hz"][j]).view(np.uint32) != w_cfo.view(np.uint32))
            if len(w_soft) >= 648:
                diff = llr[j, :648].view(np.uint32) != w_soft[:648].view(np.uint32)
                bad["soft_frames"] += int(diff.any())
                bad["soft_bits"] += int(diff.sum())
                bad["ok"] += int(int(okd[j]) != w_ok)
                bad["iters"] += int(int(iters[j]) != w_it)
                bad["info"] += int(bool(w_ok) and not np.array_equal(info[j, :21], w_info[:21]))
                decoded += w_ok
    entry = dict(frames=n, mismatches=bad, ref_decoded=int(decoded), ref_detected=int(sum(w[1] for w in want)),
                 ref_seconds=round(t_ref, 1), ref_cores=_cores())
    _report(name, entry)
    return entry


def test_c3_mcdpsk_chirp_chain_awgn_wide(ctx, ref):
    e = _mcdpsk_case(ctx, "c3_mcdpsk_dbpsk4x_chirp_awgn-8", -8.0, N_AWGN)
    assert all(v == 0 for v in e["mismatches"].values()), e
    assert e["ref_detected"] >= 0.99 * e["frames"]


@pytest.mark.parametrize("cond,cname", [(2, "moderate"), (3, "poor")])
def test_c3_mcdpsk_chirp_chain_watterson_wide(ctx, ref, cond, cname):
    e = _mcdpsk_case(ctx, f"c3_mcdpsk_dbpsk4x_chirp_watterson_{cname}_0dB", 0.0, max(128, N_FADED // 2), fade_cond=cond)
    assert all(v == 0 for v in e["mismatches"].values()), e


# ---------------------------------------------------------------------------------------------
# Synchronisers added in round 2: OFDM_COX acquisition and the rewritten OFDM data sync, wide
# ---------------------------------------------------------------------------------------------
def _cox_ref_worker(span):
    ref = Ref()
    ref.lib.ref_quiet()
    cfg, wins = _G["cox_cfg"], _G["cox_wins"]
    return [ref.ofdm_cox_search_sync(cfg, wins[i], 0.8, 0.0) for i in range(span[0], span[1])]


def _dsync_ref_worker(span):
    ref = Ref()
    ref.lib.ref_quiet()
    cfg, wins, cfos = _G["ds_cfg"], _G["ds_wins"], _G["ds_cfo"]
    out = []
    for i in range(span[0], span[1]):
        r = ref.ofdm_data_sync(cfg, wins[i], float(cfos[i]), 0.3)
        out.append((r.detected, r.start_sample, r.correlation, r.aux))
    return out


def test_ofdm_cox_search_sync_wide(ctx, ref):
    """RIA_PARITY_FADED (default 1024) windows through the batched searchForSync and the reference's: found flag, LTS
    position, CFO bits and the noise floor left behind must be identical"""
    import torch
    from ria_b200 import ofdm, sync
    n, window = N_FADED, 30000
    rng = np.random.default_rng(404)
    cfg = ModemConfig.make(QAM64, 4, 1)
    bps = cfg.data_carriers() * BITS_PER_CARRIER[cfg.modulation]
    pool = []
    for i in range(8):
        frame = ref.make_data_frame("K1ABC", "W2XYZ", i, rng.integers(0, 256, size=4 * BYTES_PER_CW[R3_4] - 19, dtype=np.uint8))
        pool.append(ref.ofdm_cox_tx_frame(cfg, ref.encode_fixed_frame(frame, R3_4, True, bps)))
    wins = np.zeros((n, window), np.float32)
    for i in range(n):
        kind = i % 10
        if kind == 9:                                           # noise only
            wins[i] = rng.standard_normal(window).astype(np.float32) * np.float32(0.05)
            continue
        tx = pool[i % 8]
        lead = int(rng.integers(0, 9000))
        seg = tx[: window - lead]
        wins[i, lead:lead + len(seg)] = seg
        p = float(np.mean(tx[1120:].astype(np.float64) ** 2))
        snr = float(rng.choice([6.0, 10.0, 14.0, 20.0, 28.0]))
        wins[i] += rng.standard_normal(window).astype(np.float32) * np.float32(np.sqrt(p / 10 ** (snr / 10)))
    _G["cox_cfg"], _G["cox_wins"] = cfg, wins
    t0 = time.time()
    want = _fan_out(_cox_ref_worker, n)
    t_ref = time.time() - t0
    rcfg = ofdm.ModemConfig.from_buffer_copy(bytes(cfg))
    nf = torch.zeros(n, device="cuda")
    got = sync.results(sync.ofdm_cox_search_sync_batch(rcfg, torch.from_numpy(wins).cuda(), 0.8, nf, ctx))
    nf = nf.cpu().numpy()
    bad = dict(found=0, position=0, cfo=0, noise_floor=0)
    n_found = 0
    for i, (f, pos, cfo, nfr) in enumerate(want):
        bad["found"] += int(bool(got["detected"][i]) != f)
        bad["noise_floor"] += int(np.float32(nfr).view(np.uint32) != nf[i].view(np.uint32))
        if f and got["detected"][i]:
            n_found += 1
            bad["position"] += int(got["start_sample"][i] != pos)
            bad["cfo"] += int(np.float32(got["cfo_hz"][i]).view(np.uint32) != np.float32(cfo).view(np.uint32))
    _report("ofdm_cox_search_sync", dict(windows=n, found=n_found, mismatches=bad, reference_seconds=round(t_ref, 1)))
    assert sum(bad.values()) == 0 and n_found > n // 3, (bad, n_found)


def test_ofdm_data_sync_wide(ctx, ref):
    """RIA_PARITY_FRAMES (default 4096) windows through the batched detectDataSync and the reference's: detection flag,
    training position and burst marker identical, correlation within 1e-5"""
    import torch
    from ria_b200 import ofdm, sync
    from tests.ofdm_common import apply_cfo
    n, window = N_AWGN, 9 * 1120
    rng = np.random.default_rng(505)
    cfg = ModemConfig.make(DQPSK, 10, 1)
    bps = cfg.data_carriers() * BITS_PER_CARRIER[cfg.modulation]
    pool = []
    for i in range(8):
        frame = ref.make_data_frame("K1ABC", "W2XYZ", i, rng.integers(0, 256, size=4 * BYTES_PER_CW[R1_2] - 19, dtype=np.uint8))
        tx = ref.ofdm_tx_frame(cfg, ref.encode_fixed_frame(frame, R1_2, True, bps))
        pool.append(tx)
        m = tx.copy(); m[:1120] = -m[:1120]                      # burst-interleave marker
        pool.append(m)
    wins = np.zeros((n, window), np.float32)
    cfos = np.zeros(n, np.float32)
    for i in range(n):
        tx = pool[i % 16]
        mode = i % 5
        if mode == 4:                                           # starts inside a burst: no quiet lead, eight-symbol search
            start = int(rng.integers(1200, 6000))
            seg = np.concatenate([pool[(i + 3) % 16][-start:], tx])[:window]
            wins[i, : len(seg)] = seg
            sc = 1.0
        else:
            lead = int(rng.integers(100, 3500))
            seg = tx[: window - lead]
            wins[i, lead:lead + len(seg)] = seg
            sc = 0.2
        p = float(np.mean(tx.astype(np.float64) ** 2))
        snr = float(rng.choice([6.0, 10.0, 16.0, 24.0]))
        wins[i] += rng.standard_normal(window).astype(np.float32) * np.float32(np.sqrt(p / 10 ** (snr / 10)) * sc)
        cfos[i] = np.float32(rng.uniform(-5, 5)) if i % 3 == 0 else np.float32(0)
    _G["ds_cfg"], _G["ds_wins"], _G["ds_cfo"] = cfg, wins, cfos
    t0 = time.time()
    want = _fan_out(_dsync_ref_worker, n)
    t_ref = time.time() - t0
    rcfg = ofdm.ModemConfig.from_buffer_copy(bytes(cfg))
    got = sync.results(sync.ofdm_data_sync_batch(rcfg, torch.from_numpy(wins).cuda(), torch.from_numpy(cfos).cuda(), 0.3, ctx))
    bad = dict(detected=0, position=0, marker=0, correlation=0)
    n_det = 0
    for i, (det, pos, corr, aux) in enumerate(want):
        bad["detected"] += int(bool(got["detected"][i]) != bool(det))
        bad["correlation"] += int(abs(got["correlation"][i] - corr) > 1e-5 * max(1.0, corr))
        if det and got["detected"][i]:
            n_det += 1
            bad["position"] += int(got["start_sample"][i] != pos)
            bad["marker"] += int(got["aux"][i] != aux)
    _report("ofdm_data_sync", dict(windows=n, detected=n_det, mismatches=bad, reference_seconds=round(t_ref, 1)))
    assert sum(bad.values()) == 0 and n_det > n // 2, (bad, n_det)


def _zc_ref_worker(span):
    ref = Ref()
    ref.lib.ref_quiet()
    z, wins, cfos = _G["zc_cfg"], _G["zc_wins"], _G["zc_cfo"]
    out = []
    for i in range(span[0], span[1]):
        r = ref.zc_detect(z, wins[i], 0.2, 4 | 8, float(cfos[i]))
        out.append((r.detected, r.start_sample, r.root, r.correlation, r.cfo_hz))
    return out


def test_zc_detect_production_window_wide(ctx, ref):
    """RIA_PARITY_FADED (default 1024) production-size windows (31 120 samples, roots DATA | CONTROL, known CFO on a third
    of them): detection flag, position, root and CFO bits identical, correlation within 1e-5"""
    import torch
    from oracle.bindings import ZcConfig
    from ria_b200 import sync
    from tests.ofdm_common import apply_cfo
    n, window = N_FADED, 31120
    rng = np.random.default_rng(606)
    z = ZcConfig.default()
    pre = {t: ref.zc_preamble(z, t) for t in (2, 3)}
    wins = np.zeros((n, window), np.float32)
    cfos = np.zeros(n, np.float32)
    for i in range(n):
        if i % 9 != 8:
            sig = pre[2 + (i % 2)]
            cfo = float(rng.uniform(-25, 25)) if i % 3 == 0 else 0.0
            if cfo:
                sig = apply_cfo(sig, cfo)
                cfos[i] = np.float32(cfo + rng.uniform(-2, 2)) if i % 2 else np.float32(0)
            pos = int(rng.integers(0, window - len(sig) - 600))
            wins[i, pos:pos + len(sig)] += sig
            tail = window - pos - len(sig)
            wins[i, pos + len(sig):] += 0.3 * np.sin(2 * np.pi * 1200 * np.arange(tail) / 48000).astype(np.float32)
        snr = float(rng.choice([-12.0, -6.0, 0.0, 6.0, 15.0]))
        wins[i] += rng.standard_normal(window).astype(np.float32) * np.float32(np.sqrt(0.32 / 10 ** (snr / 10)))
    _G["zc_cfg"], _G["zc_wins"], _G["zc_cfo"] = z, wins, cfos
    t0 = time.time()
    want = _fan_out(_zc_ref_worker, n)
    t_ref = time.time() - t0
    zs = sync.ZCSync(sync.ZCConfig.from_buffer_copy(bytes(z)), ctx)
    got = sync.results(zs.detect_batch(torch.from_numpy(wins).cuda(), 0.2, 4 | 8, torch.from_numpy(cfos).cuda()))
    bad = dict(detected=0, position=0, root=0, correlation=0, cfo=0)
    n_det = 0
    for i, (det, pos, root, corr, cfo) in enumerate(want):
        bad["detected"] += int(bool(got["detected"][i]) != bool(det))
        bad["root"] += int(got["root"][i] != root)
        bad["correlation"] += int(abs(got["correlation"][i] - corr) > 1e-5 * max(1.0, abs(corr)))
        if det and got["detected"][i]:
            n_det += 1
            bad["position"] += int(got["start_sample"][i] != pos)
            bad["cfo"] += int(np.float32(got["cfo_hz"][i]).view(np.uint32) != np.float32(cfo).view(np.uint32))
    _report("zc_detect_31120", dict(windows=n, detected=n_det, mismatches=bad, reference_seconds=round(t_ref, 1)))
    assert sum(bad.values()) == 0 and n_det > n // 2, (bad, n_det)

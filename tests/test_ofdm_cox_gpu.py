"""GPU parity of the OFDM_COX acquisition against the unmodified reference:
OFDMModulator::generatePreamble + modulate (transmit), Impl::measureCorrelation (tap) and
OFDMDemodulator::searchForSync = OFDMNvisWaveform::detectSync, then processPresynced at the position found."""
import os

import numpy as np
import pytest
import torch

from oracle.bindings import (BITS_PER_CARRIER, BYTES_PER_CW, R1_2, R3_4, DQPSK, QPSK, QAM16, QAM64, D8PSK)
from tests.ofdm_common import apply_cfo, make_cfg

pytestmark = pytest.mark.gpu

CONFIGS = [(QPSK, 5, 1, R1_2), (DQPSK, 10, 1, R1_2), (QAM64, 4, 1, R3_4), (DQPSK, 2, 0, R1_2), (D8PSK, 8, 1, R1_2),
           (QAM16, 5, 1, R3_4)]


def _ria_cfg(cfg):
    from ria_b200 import ofdm
    return ofdm.ModemConfig.from_buffer_copy(bytes(cfg))


def _coded(ref, cfg, rate, rng, seq=1):
    bpc = BYTES_PER_CW[rate]
    payload = rng.integers(0, 256, size=4 * bpc - 19, dtype=np.uint8)
    frame = ref.make_data_frame("K1ABC", "W2XYZ", seq, payload)
    bps = cfg.data_carriers() * BITS_PER_CARRIER[cfg.modulation]
    return ref.encode_fixed_frame(frame, rate, True, bps)


@pytest.mark.parametrize("mod,spacing,pilots,rate", CONFIGS)
def test_cox_tx_is_sample_identical(ctx, ref, mod, spacing, pilots, rate):
    from ria_b200 import ofdm
    cfg_o = make_cfg(mod, spacing, pilots)
    cfg = _ria_cfg(cfg_o)
    rng = np.random.default_rng(5)
    coded = rng.integers(0, 256, size=(4, 324), dtype=np.uint8)
    got = ofdm.ofdm_cox_tx_frames(cfg, torch.from_numpy(coded).cuda(), ctx).cpu().numpy()
    for i in range(len(coded)):
        want = ref.ofdm_cox_tx_frame(cfg_o, bytes(coded[i]))
        assert got.shape[1] == len(want)
        assert np.array_equal(got[i].view(np.uint32), want.view(np.uint32)), (i, np.abs(got[i] - want).max())
    assert (got[:, :1120] == 0).all()


def _windows(ref, rng, n, window, cfg_o, rate):
    """COX frames at random positions, SNRs and carrier offsets, plus noise-only, silent and DC-offset windows."""
    wins, meta = [], []
    for i in range(n):
        kind = ("sig", "sig", "sig", "noise", "sig", "silence", "sig_dc", "sig_low")[i % 8]
        w = np.zeros(window, np.float32)
        lead = -1
        if kind.startswith("sig"):
            tx = ref.ofdm_cox_tx_frame(cfg_o, _coded(ref, cfg_o, rate, rng, seq=i))
            cfo = float(rng.uniform(-30, 30)) if i % 3 else 0.0
            if cfo:
                tx = apply_cfo(tx, cfo)
            lead = int(rng.integers(0, max(1, window - 14000)))
            seg = tx[: window - lead]
            w[lead:lead + len(seg)] = seg
            snr = 3.0 if kind == "sig_low" else float(rng.choice([10, 16, 22, 30]))
            p = float(np.mean(tx[1120:].astype(np.float64) ** 2))
            w += rng.standard_normal(window).astype(np.float32) * np.float32(np.sqrt(p / 10 ** (snr / 10)))
            if kind == "sig_dc":
                w += np.float32(0.05)
        elif kind == "noise":
            w += rng.standard_normal(window).astype(np.float32) * np.float32(0.1)
        wins.append(w)
        meta.append((kind, lead))
    return wins, meta


def test_correlation_tap_is_bit_exact(ctx, ref):
    from ria_b200 import sync
    cfg_o = make_cfg(QPSK, 5, 1)
    cfg = _ria_cfg(cfg_o)
    rng = np.random.default_rng(31)
    window = 24000
    wins, meta = _windows(ref, rng, 16, window, cfg_o, R1_2)
    rows, offs = [], []
    for i, w in enumerate(wins):
        lead = meta[i][1]
        cands = [0, 64, 4000, window - 1120] + ([lead + 1120 + d for d in (-300, -8, 0, 8, 640, 1120, 2000)] if lead >= 0 else [])
        for o in cands:
            if 0 <= o and o + 1120 <= window:
                rows.append(i); offs.append(o)
    x = torch.from_numpy(np.stack([wins[i] for i in rows])).cuda()
    got = sync.ofdm_cox_correlation_batch(cfg, x, torch.tensor(offs, dtype=torch.int32, device="cuda"), ctx).cpu().numpy()
    want = np.array([ref.ofdm_cox_correlation(cfg_o, wins[i], o) for i, o in zip(rows, offs)], np.float32)
    same = got.view(np.uint32) == want.view(np.uint32)
    assert same.all(), (int((~same).sum()), got[~same][:6], want[~same][:6])
    assert (want > 0.9).sum() >= 10 and (want < 0.5).sum() >= 10


@pytest.mark.parametrize("mod,spacing,pilots,rate,window", [(QPSK, 5, 1, R1_2, 24000), (DQPSK, 10, 1, R1_2, 40000),
                                                           (QAM64, 4, 1, R3_4, 65536)])
def test_search_sync_matches_reference(ctx, ref, mod, spacing, pilots, rate, window):
    from ria_b200 import sync
    cfg_o = make_cfg(mod, spacing, pilots)
    cfg = _ria_cfg(cfg_o)
    rng = np.random.default_rng(window)
    n = 24
    wins, meta = _windows(ref, rng, n, window, cfg_o, rate)
    x = torch.from_numpy(np.stack(wins)).cuda()
    for thr in (0.8, 0.6):
        nf = torch.zeros(n, device="cuda")
        got = sync.results(sync.ofdm_cox_search_sync_batch(cfg, x, thr, nf, ctx))
        nf1 = nf.cpu().numpy().copy()
        # a second call carries the noise floor on, as a demodulator object that lives across calls does
        got2 = sync.results(sync.ofdm_cox_search_sync_batch(cfg, x.flip(0).contiguous(), thr, nf, ctx))
        nf2 = nf.cpu().numpy()
        n_found = near = 0
        for i, w in enumerate(wins):
            f, pos, cfo, nfr = ref.ofdm_cox_search_sync(cfg_o, w, thr, 0.0)
            g = got[i]
            assert bool(g["detected"]) == f, (i, meta[i], g, pos)
            assert np.float32(nfr).view(np.uint32) == nf1[i].view(np.uint32), (i, meta[i], nfr, nf1[i])
            if f:
                assert g["start_sample"] == pos, (i, meta[i], g["start_sample"], pos)
                assert np.float32(g["cfo_hz"]).view(np.uint32) == np.float32(cfo).view(np.uint32), (i, g["cfo_hz"], cfo)
                assert g["correlation"] == np.float32(0.9)
                n_found += 1
                near += abs(pos - (meta[i][1] + 1120 + 4480)) <= 8
            # second call: window n-1-i with the floor left by window i of the first call
            j = n - 1 - i
            f2, pos2, cfo2, nfr2 = ref.ofdm_cox_search_sync(cfg_o, wins[j], thr, float(nf1[i]))
            g2 = got2[i]
            assert bool(g2["detected"]) == f2, (i, j, meta[j])
            assert np.float32(nfr2).view(np.uint32) == nf2[i].view(np.uint32), (i, j)
            if f2:
                assert g2["start_sample"] == pos2 and np.float32(g2["cfo_hz"]).view(np.uint32) == np.float32(cfo2).view(np.uint32)
        assert n_found >= 6 and near >= n_found // 2, (n_found, near)


def test_edges(ctx, ref):
    from ria_b200 import sync
    cfg_o = make_cfg(QPSK, 5, 1)
    cfg = _ria_cfg(cfg_o)
    for window in (3000, 8000, 8960):
        got = sync.results(sync.ofdm_cox_search_sync_batch(cfg, torch.randn((3, window), device="cuda"), 0.8, None, ctx))
        assert (got["detected"] == 0).all()
        f, pos, cfo, _ = ref.ofdm_cox_search_sync(cfg_o, np.random.default_rng(1).standard_normal(window).astype(np.float32), 0.8)
        assert not f
    with pytest.raises(Exception):
        sync.ofdm_cox_search_sync_batch(cfg, torch.zeros((1, 70000), device="cuda"), 0.8, None, ctx)
    empty = sync.ofdm_cox_search_sync_batch(cfg, torch.zeros((0, 20000), device="cuda"), 0.8, None, ctx)
    assert empty.shape[0] == 0


def test_acquire_then_demodulate_like_the_cox_waveform(ctx, ref):
    """OFDMNvisWaveform::detectSync + process (ofdm_cox_waveform.cpp:121-218) for a batch: search, then the presynced
    demodulator at the LTS position with the CFO found and the initial phase the waveform derives from them."""
    from ria_b200 import ofdm, sync
    cfg_o = make_cfg(QPSK, 5, 1)
    cfg = _ria_cfg(cfg_o)
    rng = np.random.default_rng(9)
    n, window = 12, 48000
    wins = []
    for i in range(n):
        tx = ref.ofdm_cox_tx_frame(cfg_o, _coded(ref, cfg_o, R1_2, rng, seq=i))
        cfo = float(rng.uniform(-12, 12)) if i % 2 else 0.0
        if cfo:
            tx = apply_cfo(tx, cfo)
        w = np.zeros(window, np.float32)
        lead = int(rng.integers(100, 6000))
        w[lead:lead + len(tx)] = tx
        p = float(np.mean(tx[1120:].astype(np.float64) ** 2))
        w += rng.standard_normal(window).astype(np.float32) * np.float32(np.sqrt(p / 10 ** (22 / 10)))
        wins.append(w)
    x = torch.from_numpy(np.stack(wins)).cuda()
    res = sync.results(sync.ofdm_cox_search_sync_batch(cfg, x, 0.8, None, ctx))
    assert res["detected"].all()
    import ctypes
    import ria_b200
    flen = ria_b200.lib().ria_ofdm_tx_frame_samples(ctypes.addressof(cfg), 324)
    starts = res["start_sample"].astype(np.int64)
    cfos = res["cfo_hz"].astype(np.float32)
    # initial phase as OFDMNvisWaveform::process computes it: a double expression (M_PI) rounded to float, then
    # wrapped to [-pi, pi] with double arithmetic rounded to float at every step (ofdm_cox_waveform.cpp:172-176)
    wrapped = np.zeros(n, np.float32)
    for i in range(n):
        r = ref.ofdm_cox_search_sync(cfg_o, wins[i], 0.8, 0.0)
        assert r[1] == starts[i] and np.float32(r[2]) == cfos[i]
        p = np.float32(-2.0 * np.pi * float(cfos[i]) * float(starts[i]) / 48000.0)
        while float(p) > np.pi:
            p = np.float32(float(p) - 2.0 * np.pi)
        while float(p) < -np.pi:
            p = np.float32(float(p) + 2.0 * np.pi)
        wrapped[i] = p
    idx = torch.from_numpy(starts).cuda()[:, None] + torch.arange(flen, device="cuda")[None, :]
    frames = torch.gather(x, 1, idx)
    dem = ofdm.OFDMDemodulator(cfg, ctx)
    out = dem.process_presynced_batch(frames, torch.from_numpy(cfos).cuda(), torch.from_numpy(wrapped).cuda())
    llr = out["llr"].cpu().numpy(); n_llr = out["n_llr"].cpu().numpy()
    for i in range(n):
        r = ref.ofdm_process_presynced(cfg_o, wins[i][starts[i]:starts[i] + flen], float(cfos[i]), float(wrapped[i]))
        assert r["ready"] and n_llr[i] == len(r["soft"])
        assert np.array_equal(llr[i, :n_llr[i]].view(np.uint32), r["soft"].view(np.uint32)), i


def test_cox_chain_one_call_matches_the_stages(ctx, ref):
    """ria_ofdm_cox_rx_frames_dev / _host: search + process + complete frame decode in one call; the decoded frames are the
    ones that were sent, the sync results those of the search alone, windows without a frame come back not valid"""
    import ria_b200
    from ria_b200 import ofdm, sync
    cfg_o = make_cfg(QAM64, 4, 1)
    cfg = _ria_cfg(cfg_o)
    rng = np.random.default_rng(77)
    n, window, frame_len = 20, 26000, 12 * 1120
    bps = cfg_o.data_carriers() * BITS_PER_CARRIER[cfg_o.modulation]
    wins, sent = [], []
    for i in range(n):
        w = np.zeros(window, np.float32)
        frame = None
        if i % 5 != 4:
            frame = ref.make_data_frame("K1ABC", "W2XYZ", i, rng.integers(0, 256, size=4 * BYTES_PER_CW[R3_4] - 19, dtype=np.uint8))
            tx = ref.ofdm_cox_tx_frame(cfg_o, ref.encode_fixed_frame(frame, R3_4, True, bps))
            # QAM64 survives almost no residual CFO in the reference itself: the CFO windows check the sync fields and the
            # bit-exact failure, the others the decoded frames
            cfo = float(rng.uniform(-1.5, 1.5)) if i % 4 == 1 else 0.0
            lead = int(rng.integers(0, 5000))
            if cfo:
                tx = apply_cfo(np.concatenate([np.zeros(lead, np.float32), tx]), cfo)[lead:]
            w[lead:lead + len(tx)] = tx
            p = float(np.mean(tx[1120:].astype(np.float64) ** 2))
            w += rng.standard_normal(window).astype(np.float32) * np.float32(np.sqrt(p / 10 ** (30.0 / 10)))
        else:
            w += rng.standard_normal(window).astype(np.float32) * np.float32(0.05)
        wins.append(w); sent.append(frame)
    x = np.stack(wins)
    ctx.set_decode_flags(ria_b200.DECODE_FULL)
    chain = ofdm.OfdmCoxRxChain(cfg, R3_4, True, ctx)
    data, status, snr, sy = chain.process_windows(torch.from_numpy(x).cuda(), frame_len, 0.8)
    torch.cuda.synchronize()
    data = data.cpu().numpy(); st = ofdm.status_array(status); sy = sync.results(sy)
    alone = sync.results(sync.ofdm_cox_search_sync_batch(cfg, torch.from_numpy(x).cuda(), 0.8, None, ctx))
    h_data, h_st, h_snr, h_sy = chain.process_windows_host(x, frame_len, 0.8)
    ctx.set_decode_flags(0)
    n_ok = n_repaired_wrong = 0
    for i in range(n):
        assert sy["detected"][i] == alone["detected"][i] and sy["start_sample"][i] == alone["start_sample"][i]
        assert np.float32(sy["cfo_hz"][i]).view(np.uint32) == np.float32(alone["cfo_hz"][i]).view(np.uint32)
        f, pos, cfo, _ = ref.ofdm_cox_search_sync(cfg_o, wins[i], 0.8, 0.0)
        assert bool(sy["detected"][i]) == f and (not f or sy["start_sample"][i] == pos)
        ok = bool(st["all_ok"][i]) and bool(st["header_valid"][i]) and bool(st["frame_crc_ok"][i])
        if f and pos + frame_len <= window:
            # the reference's own process + decodeFixedFrame at the position / CFO / phase the waveform would use
            ph = np.float32(-2.0 * np.pi * float(cfo) * float(pos) / 48000.0)
            while float(ph) > np.pi:
                ph = np.float32(float(ph) - 2.0 * np.pi)
            while float(ph) < -np.pi:
                ph = np.float32(float(ph) + 2.0 * np.pi)
            r = ref.ofdm_process_presynced(cfg_o, wins[i][pos:pos + frame_len], float(cfo), float(ph))
            rd, rok = ref.decode_fixed_frame_full(r["soft"], R3_4, True, bps)
            rst = ref.frame_status_reassembled(rd, rok, BYTES_PER_CW[R3_4])
            assert ok == bool(rok.all() and rst.header_valid and rst.frame_crc_ok), (i, ok, rok)
            assert np.array_equal(data[i], np.frombuffer(bytes(rd), np.uint8)[: data.shape[1]]), i
        if sent[i] is None:
            assert not ok
        elif ok:
            # decodeFixedFrame's CRC-guided bit-flip search (frame_v2.cpp:1559-1874, up to 4 suspect bits against a 16-bit
            # CRC) can settle on a wrong frame whose CRC matches; the reference does the same (asserted above) and it must
            # stay the exception
            if bytes(data[i, : len(sent[i])]) == sent[i]:
                n_ok += 1
            else:
                n_repaired_wrong += 1
        assert np.array_equal(h_data[i], data[i]) and h_st[i] == st[i] and h_sy["start_sample"][i] == sy["start_sample"][i]
    assert n_ok >= 10 and n_repaired_wrong <= 2, (n_ok, n_repaired_wrong)


GOLD_COX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cox_golden.npz")


@pytest.mark.parametrize("name,mod,spacing", [("qam64_sp4", QAM64, 4), ("dqpsk_sp5", DQPSK, 5)])
def test_search_sync_against_committed_golden(ctx, name, mod, spacing):
    """tests/golden/cox_golden.npz: outputs of the unmodified reference's searchForSync / measureCorrelation on seeded
    windows (tests/golden/make_golden.py cox); runs without oracle/_ref."""
    from ria_b200 import sync
    g = np.load(GOLD_COX)
    cfg = _ria_cfg(make_cfg(mod, spacing, 1))
    x = torch.from_numpy(g[f"{name}_win"].astype(np.float32)).cuda()
    for tag in ("a", "b"):
        nf = torch.full((x.shape[0],), float(g[f"{name}_{tag}_nf_in"]), device="cuda")
        got = sync.results(sync.ofdm_cox_search_sync_batch(cfg, x, float(g[f"{name}_{tag}_thr"]), nf, ctx))
        assert np.array_equal(got["detected"].astype(np.uint8), g[f"{name}_{tag}_found"])
        assert np.array_equal(nf.cpu().numpy().view(np.uint32), g[f"{name}_{tag}_nf_out"].view(np.uint32))
        hit = g[f"{name}_{tag}_found"].astype(bool)
        assert np.array_equal(got["start_sample"][hit].astype(np.int64), g[f"{name}_{tag}_pos"][hit])
        assert np.array_equal(got["cfo_hz"][hit].astype(np.float32).view(np.uint32), g[f"{name}_{tag}_cfo"][hit].view(np.uint32))
    corr = sync.ofdm_cox_correlation_batch(cfg, x, torch.from_numpy(g[f"{name}_corr_off"]).cuda(), ctx).cpu().numpy()
    assert np.array_equal(corr.view(np.uint32), g[f"{name}_corr"].view(np.uint32))

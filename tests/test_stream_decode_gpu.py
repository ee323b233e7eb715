"""GPU: batched StreamingDecoder::decodeMCDPSKFrame (ria_b200.mcdpsk.McdpskFrameDecoder) against the
reference's own StreamingDecoder object (oracle/_ref, private member reached through the shim) on identical soft
bits: control frames, multi-codeword data frames, receptions too noisy to decode, truncated buffers, and HARQ
retransmissions that only decode after chase combining (streaming_decoder.cpp:2762-2789)."""
import numpy as np
import pytest

from oracle.bindings import R1_4, R1_2

pytestmark = pytest.mark.gpu


def _soft(cws, snr_db, rng, slots):
    """BPSK soft bits of the coded bytes [n_cw, 81] at Es/N0 snr_db, padded with noise-only codeword slots"""
    bits = np.unpackbits(cws, axis=1)[:, :648].astype(np.float32)
    snr = 10 ** (snr_db / 10)
    s = 1.0 - 2.0 * bits
    llr = 2.0 * (s + rng.standard_normal(s.shape).astype(np.float32) / np.sqrt(snr)) * snr
    out = (rng.standard_normal((slots, 648)) * 0.5).astype(np.float32)
    out[: len(llr)] = llr[:slots]
    return out.reshape(-1)


def _compare(ref, h, dec, soft_rows, rate):
    import torch
    got = dec.decode_batch(torch.from_numpy(np.stack(soft_rows)).cuda())
    n_ok = 0
    for i, row in enumerate(soft_rows):
        res, data = ref.stream_decode_mcdpsk_frame(h, row, rate)
        assert got["success"][i] == res.success, (i, got["success"][i], res.success)
        assert got["codewords_ok"][i] == res.codewords_ok and got["codewords_failed"][i] == res.codewords_failed, \
            (i, got["codewords_ok"][i], res.codewords_ok, got["codewords_failed"][i], res.codewords_failed)
        if res.success or res.codewords_ok:
            assert got["frame_type"][i] == res.frame_type, i
        assert got["frame_len"][i] == len(data), (i, got["frame_len"][i], len(data))
        assert bytes(got["frame"][i, : len(data)]) == data, i
        n_ok += res.success
    return n_ok, got


@pytest.mark.parametrize("rate,snr_db", [(R1_4, -1.0), (R1_2, 2.2)])
def test_decode_mcdpsk_frame_matches_streaming_decoder(ctx, ref, rate, snr_db):
    from ria_b200 import fec, mcdpsk
    rng = np.random.default_rng(40 + rate)
    slots = 6
    rows = []
    for i in range(48):
        kind = i % 6
        payload = rng.integers(0, 256, size=int(rng.integers(1, 60)), dtype=np.uint8)
        frame = ref.make_data_frame("K1ABC", "W2XYZ", 100 + i, payload)          # distinct seq: no cache interaction
        cws = ref.encode_frame_with_ldpc(frame, rate)
        snr = snr_db + (0.0, 1.5, 3.0, -2.5, 6.0, 0.5)[kind]
        avail = slots if kind != 4 else max(1, len(cws) - 1)                      # kind 4: buffer ends before the last codeword
        row = _soft(cws, snr, rng, slots)
        if avail < slots:
            row = row.copy()
            row[avail * 648:] = 0.0
        rows.append(row)
    # pure noise and a corrupted magic
    rows.append((rng.standard_normal(slots * 648) * 2).astype(np.float32))
    h = ref.stream_decoder()
    try:
        dec = mcdpsk.McdpskFrameDecoder(rate, ctx, fec.ChaseCache(max_entries=64, ctx=ctx))
        n_ok, _ = _compare(ref, h, dec, rows, rate)
    finally:
        ref.stream_decoder_free(h)
    assert 8 <= n_ok < len(rows)


def test_harq_retransmissions_decode_after_chase_combining(ctx, ref):
    """Receptions of the same frames (same seq / src / dst) too noisy for their data codewords: the first round
    leaves failures in the chase cache, later rounds decode the combined soft bits -- round by round the results
    must equal the reference decoder's, whose own ChaseCache sees the same sequence of receptions."""
    from ria_b200 import fec, mcdpsk
    rng = np.random.default_rng(7)
    slots, n_frames = 4, 12                                      # 12 keys <= 16 cache entries: no evictions on either side
    frames = [ref.encode_frame_with_ldpc(ref.make_data_frame("N0CALL", "W1AW", 300 + i, rng.integers(0, 256, size=40, dtype=np.uint8)), R1_4)
              for i in range(n_frames)]
    h = ref.stream_decoder()
    try:
        dec = mcdpsk.McdpskFrameDecoder(R1_4, ctx, fec.ChaseCache(max_entries=16, ctx=ctx))
        ok_per_round = []
        for rnd in range(4):
            rows = []
            for cws in frames:
                row = _soft(cws, -4.6, rng, slots).reshape(slots, 648)
                row[0] = _soft(cws[:1], 2.0, rng, 1)             # the header codeword always gets through
                rows.append(row.reshape(-1))
            n_ok, _ = _compare(ref, h, dec, rows, R1_4)
            ok_per_round.append(n_ok)
    finally:
        ref.stream_decoder_free(h)
    assert ok_per_round[0] < n_frames, ok_per_round               # the first reception alone is not enough ...
    assert max(ok_per_round[1:]) > ok_per_round[0], ok_per_round  # ... combining recovers frames
    assert dec.stats["chase_recoveries"] > 0


def test_zc_acquired_chain_matches_reference_calls(ctx, ref):
    """configs[2] variant (ii): [noise][ZC DATA preamble][training][reference][multi-codeword data] in the 31 120-sample
    search window -> detectDataSync (ZC) -> process at the detected start with the detected CFO -> decodeMCDPSKFrame,
    against the same three reference calls on the same rows."""
    import torch
    from oracle.bindings import McdpskConfig, ZcConfig
    from ria_b200 import fec, mcdpsk
    from tests.ofdm_common import awgn
    rng = np.random.default_rng(11)
    cfg = McdpskConfig.make(1, 2, 10)                                   # DBPSK, 2x spreading, 10 carriers
    rcfg = mcdpsk.MultiCarrierDPSKConfig.from_buffer_copy(bytes(cfg))
    zc = ZcConfig.default()
    pre = ref.zc_preamble(zc, 2)                                        # DATA root
    window, n = 31120, 24
    rows, frame_len = [], None
    for i in range(n):
        frame = ref.make_data_frame("K1ABC", "W2XYZ", 500 + i, rng.integers(0, 256, size=30, dtype=np.uint8))
        cws = ref.encode_frame_with_ldpc(frame, R1_4)                   # 3 codewords
        body = ref.mcdpsk_tx_frame(cfg, cws.reshape(-1))
        frame_len = len(body)
        lead = int(rng.integers(200, 6000))
        tx = np.concatenate([np.zeros(lead, np.float32), pre, body, np.zeros(9000 - lead, np.float32)])
        rows.append(awgn(tx, (-3.0, 0.0, 3.0)[i % 3], rng))
    rows = np.stack(rows)
    chain = mcdpsk.McdpskZcRxChain(rcfg, R1_4, ctx, fec.ChaseCache(max_entries=64, ctx=ctx), threshold=0.2)
    got, sync_t = chain.process_batch(torch.from_numpy(rows).cuda(), window, frame_len)
    from ria_b200 import sync as rsync
    sy = rsync.results(sync_t)
    h = ref.stream_decoder()
    n_ok = 0
    try:
        for i in range(n):
            s = ref.zc_detect(zc, rows[i, :window], 0.2, 4 | 8, 0.0)
            assert bool(sy["detected"][i]) == bool(s.detected), i
            if not s.detected:
                assert got["success"][i] == 0
                continue
            assert sy["start_sample"][i] == s.start_sample and np.float32(sy["cfo_hz"][i]) == np.float32(s.cfo_hz), i
            r = ref.mcdpsk_process(cfg, rows[i, s.start_sample: s.start_sample + frame_len], float(s.cfo_hz))
            slots = len(r["soft"]) // 648
            res, data = ref.stream_decode_mcdpsk_frame(h, r["soft"][: slots * 648], R1_4)
            assert got["success"][i] == res.success and got["codewords_ok"][i] == res.codewords_ok, i
            assert got["frame_len"][i] == len(data) and bytes(got["frame"][i, : len(data)]) == data, i
            n_ok += res.success
    finally:
        ref.stream_decoder_free(h)
    assert n_ok >= n // 2


@pytest.mark.parametrize("mod_name,rate,spacing,esn0", [("QAM16", 3, 5, 5.4), ("DQPSK", 2, 10, 1.9), ("DQPSK", 0, 10, -2.2)])
def test_decode_frame_ofdm_matches_streaming_decoder(ctx, ref, mod_name, rate, spacing, esn0):
    """Batched StreamingDecoder::decodeFrame for OFDM receivers (ria_b200.ofdm.OfdmFrameDecoder): four-codeword data frames
    through the frame-interleaved decode, one-codeword control frames through the R1/4 fast path / raw codeword 0, the
    salvage of control frames after a failed four-codeword decode, frames too noisy to decode, and short buffers."""
    import torch
    import oracle.bindings as ob
    from oracle.bindings import BITS_PER_CARRIER, BYTES_PER_CW, ModemConfig
    from ria_b200 import ofdm
    mod = getattr(ob, mod_name)
    cfg = ModemConfig.make(mod, spacing, 1)
    bps = cfg.data_carriers() * BITS_PER_CARRIER[mod]
    bpc = BYTES_PER_CW[rate]
    rng = np.random.default_rng(900 + rate)
    snr = 10 ** (esn0 / 10)
    rows = []
    for i in range(60):
        kind = i % 6
        noise_db = (0.0, 1.0, 3.0, -2.0, 4.0, -6.0)[kind]
        if kind in (0, 1, 3, 5):                                   # four-codeword data frame, frame + channel interleaved
            frame = ref.make_data_frame("K1ABC", "W2XYZ", i, rng.integers(0, 256, size=4 * bpc - 19 - int(rng.integers(0, 30)), dtype=np.uint8))
            coded = np.frombuffer(bytes(ref.encode_fixed_frame(frame, rate, True, bps)), np.uint8)
            bits = np.unpackbits(coded)[:2592].astype(np.float32)
        else:                                                      # one-codeword control frame: R1/4, not interleaved, padded with noise
            ctl = bytearray(20)
            ctl[0:2] = b"\x55\x4c"; ctl[2] = (0x20, 0x16)[kind == 4]; ctl[4:6] = int(i).to_bytes(2, "big")
            ctl[6:12] = bytes(rng.integers(0, 256, size=6, dtype=np.uint8))
            ctl[18:20] = int(ref.crc16(bytes(ctl[:18]))).to_bytes(2, "big")
            cw = ref.ldpc_encode(0, np.frombuffer(bytes(ctl), np.uint8))[:81]
            bits = np.concatenate([np.unpackbits(cw)[:648], rng.integers(0, 2, size=2592 - 648)]).astype(np.float32)
        s = 1.0 - 2.0 * bits
        g = snr * 10 ** (noise_db / 10)
        llr = (2.0 * (s + rng.standard_normal(s.shape).astype(np.float32) / np.sqrt(g)) * g).astype(np.float32)
        if kind == 2 and i % 12 == 2:
            llr[648:] = 0.0
        rows.append(llr)
    rows = np.stack(rows)
    dec = ofdm.OfdmFrameDecoder(mod, rate, cfg.data_carriers(), True, True, ctx)
    got = dec.decode_batch(torch.from_numpy(rows).cuda())
    short = dec.decode_batch(torch.from_numpy(rows[:12, :1300]).cuda())      # buffers that end inside codeword 2
    h = ref.stream_decoder()
    n_ok = n_ctl = 0
    try:
        for tag, out, data_rows in (("full", got, rows), ("short", short, rows[:12, :1300])):
            for i, row in enumerate(data_rows):
                res, data = ref.stream_decode_ofdm_frame(h, row, True, mod, rate, cfg.data_carriers(), True)
                assert out["success"][i] == res.success, (tag, i, out["success"][i], res.success)
                assert out["codewords_ok"][i] == res.codewords_ok and out["codewords_failed"][i] == res.codewords_failed, \
                    (tag, i, out["codewords_ok"][i], res.codewords_ok, out["codewords_failed"][i], res.codewords_failed)
                assert out["frame_len"][i] == len(data) and bytes(out["frame"][i, : len(data)]) == data, (tag, i)
                if res.success:
                    assert out["frame_type"][i] == res.frame_type, (tag, i)
                    n_ok += 1
                    n_ctl += int(res.codewords_ok == 1)
    finally:
        ref.stream_decoder_free(h)
    assert n_ok >= 20 and n_ctl >= 5, (n_ok, n_ctl)

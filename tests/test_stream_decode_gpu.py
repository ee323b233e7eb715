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

"""GPU parity of one step of the receive state machine against the reference's own
StreamingDecoder::decodeCurrentFrame (src/gui/modem/streaming_decoder.cpp:1060-2125), executed through
oracle/ref_shim.cpp on a StreamingDecoder object whose ring buffer holds the same window."""
import numpy as np
import pytest
import torch

from oracle.bindings import BYTES_PER_CW, R1_2, R1_4, R3_4, DBPSK, DQPSK, QPSK, QAM16, QAM64
from tests.ofdm_common import apply_cfo

pytestmark = pytest.mark.gpu

L = 36000


def _noisy(x, snr_db, rng, ref_power):
    sigma = np.sqrt(ref_power / 10 ** (snr_db / 10))
    return (x + rng.standard_normal(len(x)).astype(np.float32) * np.float32(sigma)).astype(np.float32)


def _receptions(ref, rng, modulation, rate, n):
    """windows with control frames, data frames (some badly timed, some too noisy), noise"""
    wins, meta = [], []
    for i in range(n):
        kind = ("ack", "data", "data_off", "nack", "data_low", "noise", "data_cfo", "short")[i % 8]
        pos = int(rng.integers(200, 1500))
        w = np.zeros(L, np.float32)
        cfo = 0.0
        if kind == "noise":
            w = (rng.standard_normal(L) * 0.05).astype(np.float32)
        else:
            if kind in ("ack", "nack"):
                frame = ref.make_ack_frame("K1ABC", "W2XYZ", i, nack=(kind == "nack"))
            elif kind == "short":                                              # a two-codeword data frame
                frame = ref.make_data_frame("K1ABC", "W2XYZ", i, rng.integers(0, 256, size=BYTES_PER_CW[rate], dtype=np.uint8))
            else:
                frame = ref.make_data_frame("K1ABC", "W2XYZ", i, rng.integers(0, 256, size=4 * BYTES_PER_CW[rate] - 19, dtype=np.uint8))
            tx = ref.stream_encode(1, modulation, rate, 1, frame)
            if kind == "data_cfo":
                cfo = 1.7
                tx = apply_cfo(np.concatenate([np.zeros(pos, np.float32), tx]), cfo)[pos:]
            seg = tx[: L - pos]
            w[pos:pos + len(seg)] = seg
            p = float(np.mean(tx.astype(np.float64) ** 2))
            snr = {"data_low": 4.0}.get(kind, 30.0 if modulation in (QAM64, QAM16) else 18.0)
            w = _noisy(w, snr, rng, p)
        sync = pos + (int(rng.choice([-16, -8, 8, 16])) if kind == "data_off" else 0)
        wins.append(w)
        meta.append(dict(kind=kind, pos=pos, sync=sync, cfo=cfo))
    return wins, meta


@pytest.mark.parametrize("modulation,rate", [(QAM64, R3_4), (DQPSK, R1_2), (QPSK, R1_2), (DQPSK, R1_4), (QAM16, R3_4)])
def test_connected_ofdm_step_matches_the_reference_state_machine(ctx, ref, modulation, rate):
    from ria_b200 import stream
    rng = np.random.default_rng(100 * modulation + rate)
    n = 24
    wins, meta = _receptions(ref, rng, modulation, rate, n)
    h = ref.stream_decoder()
    ref.stream_setup_ofdm(h, True, modulation, rate)
    stepper = stream.OfdmConnectedStep(modulation, rate, ctx)
    x = torch.from_numpy(np.stack(wins)).cuda()
    sync = np.array([m["sync"] for m in meta], np.int64)
    sync_cfo = np.array([m["cfo"] for m in meta], np.float32)
    last_cfo = sync_cfo.copy()
    pending = np.zeros(n, np.int32)
    active = np.ones(n, bool)
    seen = dict(control=0, escalated=0, decoded=0, failed=0, recovered=0)
    for it in range(4):
        idx = np.nonzero(active)[0]
        if len(idx) == 0:
            break
        got = stepper.step(x[torch.from_numpy(idx).cuda()], sync[idx], sync_cfo[idx], last_cfo[idx], pending[idx])
        for j, i in enumerate(idx):
            res, data = ref.stream_step(h, wins[i], int(sync[i]), float(sync_cfo[i]), 15.0, int(pending[i]), float(last_cfo[i]))
            tag = (it, i, meta[i])
            assert got["state"][j] == res.state, (tag, got["state"][j], res.state, got["pending_total_cw"][j], res.pending_total_cw)
            if res.state == 1:
                assert got["pending_total_cw"][j] == res.pending_total_cw, tag
                seen["escalated"] += 1
            assert bool(got["has_frame"][j]) == bool(res.has_frame), (tag, got["has_frame"][j], res.has_frame)
            assert np.float32(got["last_cfo"][j]).view(np.uint32) == np.float32(res.last_cfo).view(np.uint32), (tag, got["last_cfo"][j], res.last_cfo)
            if res.has_frame:
                f = res.frame
                assert (got["success"][j], got["codewords_ok"][j], got["codewords_failed"][j]) == (f.success, f.codewords_ok, f.codewords_failed), tag
                if f.success:
                    assert got["frame_type"][j] == f.frame_type, tag
                    assert got["frame_len"][j] == f.n_bytes and bytes(got["frame"][j, : f.n_bytes]) == data, tag
                    seen["control" if f.frame_type in (0x20, 0x21) else "decoded"] += 1
                    if got["sync_pos"][j] != sync[i]:
                        seen["recovered"] += 1
                        assert got["sync_pos"][j] == res.sync_pos, tag
            if res.state == 0 and not (res.has_frame and res.frame.success):
                seen["failed"] += 1
            pending[i] = res.pending_total_cw
            active[i] = res.state == 1
    ref.stream_decoder_free(h)
    assert seen["control"] >= 4 and seen["decoded"] >= 6 and seen["failed"] >= 2, seen
    if not (modulation == DQPSK and rate == R1_4):
        assert seen["escalated"] >= 6, seen


def test_ping_energy_matches_the_reference_decision(ctx, ref):
    """a disconnected MC-DPSK receiver: chirp-only transmissions are PINGs, anything with data behind the preamble is
    not; the decision is the reference's own (decodeCurrentFrame returns a PING result)"""
    from ria_b200 import stream
    rng = np.random.default_rng(7)
    h = ref.stream_decoder()
    ref.stream_setup_mcdpsk(h, False, 10, DBPSK, R1_4, 0)
    ctl = ref.stream_min_control_samples(h)
    ping = ref.stream_encode(2, DBPSK, R1_4, 2, b"", 10, 0)
    frame = ref.make_data_frame("K1ABC", "W2XYZ", 1, rng.integers(0, 256, size=40, dtype=np.uint8))
    data = ref.stream_encode(2, DBPSK, R1_4, 0, frame, 10, 0)
    pre = 57600                                                                # dual chirp: the sync position is behind it
    rows, want = [], []
    for i in range(16):
        tx = ping if i % 2 == 0 else data
        snr = float(rng.choice([4.0, 6.0, 10.0, 15.0, 20.0] if i % 2 == 0 else [-3.0, 3.0, 10.0]))
        x = np.concatenate([tx, np.zeros(max(0, pre + ctl + 6000 - len(tx)), np.float32)])
        x = _noisy(x, snr, rng, float(np.mean(tx.astype(np.float64) ** 2)))
        seg = x[pre:pre + ctl + 6000]
        res, _ = ref.stream_step(h, seg, 0, 0.0, 10.0, 0, 0.0)
        want.append(bool(res.has_frame and res.frame.is_ping))
        rows.append(seg[:ctl])                                                 # frame_buffer of the first pass
    got = stream.ping_energy_batch(torch.from_numpy(np.stack(rows)).cuda(), 4608, ctx)
    assert list(got["is_ping"].astype(bool)) == want, (got, want)
    assert sum(want) >= 4 and sum(not w for w in want) >= 8, want
    # the RMS values themselves against the loops of :1143-1160 restated in float32
    for i, r in enumerate(rows):
        acc = np.float32(0)
        for v in r[:4608]:
            acc = np.float32(acc + np.float32(v * v))
        tr = np.sqrt(np.float32(acc / np.float32(4608)), dtype=np.float32)
        assert got["training_rms"][i].view(np.uint32) == np.float32(tr).view(np.uint32), i
    ref.stream_decoder_free(h)


@pytest.mark.parametrize("modulation,rate", [(QAM16, R3_4), (DQPSK, R1_2)])
def test_burst_group_matches_the_reference(ctx, ref, modulation, rate):
    """burst-interleaved groups: marker frame, three continuation blocks, de-interleave, four decodes; plus a group whose
    third block lost its energy (aborted) and one whose last block never arrives"""
    from ria_b200 import stream
    rng = np.random.default_rng(11 + modulation)
    h = ref.stream_decoder()
    ref.stream_setup_ofdm(h, True, modulation, rate)
    stepper = stream.OfdmConnectedStep(modulation, rate, ctx)
    block = stepper.samples_for_cw(4)
    Lw = 4 * block + 4000
    wins, sync, cfos, kinds, sent = [], [], [], [], []
    for i in range(8):
        kind = ("ok", "ok", "lost", "ok_cfo", "late", "ok_low", "ok", "ok")[i]
        frames = [ref.make_data_frame("K1ABC", "W2XYZ", 4 * i + k, rng.integers(0, 256, size=4 * BYTES_PER_CW[rate] - 19, dtype=np.uint8))
                  for k in range(4)]
        tx = ref.stream_encode_burst(modulation, rate, frames, 4)
        assert len(tx) == 4 * block
        pos = int(rng.integers(100, 2500))
        cfo = 1.3 if kind == "ok_cfo" else 0.0
        w = np.zeros(Lw, np.float32)
        w[pos:pos + len(tx)] = tx
        if cfo:
            w = apply_cfo(w, cfo)
        if kind == "lost":
            w[pos + 2 * block: pos + 3 * block] = 0.0
        p = float(np.mean(tx.astype(np.float64) ** 2))
        w = _noisy(w, 12.0 if kind == "ok_low" else 28.0, rng, p)
        if kind == "lost":
            w[pos + 2 * block: pos + 3 * block] *= np.float32(0.01)
        if kind == "late":
            w = w[: pos + 3 * block + 500].copy()
            w = np.concatenate([w, np.zeros(Lw - len(w), np.float32)])         # the batch is rectangular; see n_valid below
        wins.append(w); sync.append(pos); cfos.append(cfo); kinds.append(kind); sent.append(frames)
    x = torch.from_numpy(np.stack(wins)).cuda()
    # the late group: hand the reference only the samples that arrived; the batched call sees a window that ends there too
    late = kinds.index("late")
    n_valid = sync[late] + 3 * block + 500
    got_main, last_main = stepper.burst_group(x, np.array(sync), np.array(cfos, np.float32), np.array(cfos, np.float32), 4)
    got_late, last_late = stepper.burst_group(x[late:late + 1, :n_valid].contiguous(), np.array(sync[late:late + 1]),
                                              np.array(cfos[late:late + 1], np.float32), np.array(cfos[late:late + 1], np.float32), 4)
    n_ok = 0
    for i in range(8):
        samples = wins[i][:n_valid] if i == late else wins[i]
        res, lc = ref.stream_burst_group(h, samples, sync[i], cfos[i], cfos[i], 4)
        got, last, j = (got_late, last_late, 0) if i == late else (got_main, last_main, i)
        queued = [k for k in range(4) if got[k]["queued"][j]]
        if res is None:                                                        # still accumulating: nothing delivered
            assert kinds[i] == "late" and queued == []
            continue
        assert np.float32(last[j]).view(np.uint32) == np.float32(lc).view(np.uint32), (i, kinds[i], last[j], lc)
        assert len(queued) == len(res), (i, kinds[i], queued, len(res))
        for k, (r, data) in zip(queued, res):
            g = got[k]
            assert (g["success"][j], g["codewords_ok"][j], g["codewords_failed"][j]) == (r.success, r.codewords_ok, r.codewords_failed), (i, k)
            if r.success:
                assert bytes(g["frame"][j, : r.n_bytes]) == data == sent[i][k], (i, k)
                n_ok += 1
        if kinds[i] == "lost":
            assert res == []
    ref.stream_decoder_free(h)
    assert n_ok >= 16, n_ok


@pytest.mark.parametrize("connected,modulation,rate,spread", [(False, DBPSK, R1_4, 0), (True, DQPSK, R1_2, 0), (False, DBPSK, R1_4, 1)])
def test_mcdpsk_step_matches_the_reference_state_machine(ctx, ref, connected, modulation, rate, spread):
    """MC-DPSK receivers: PING classification, the codeword-0 peek with the handshake rules (a disconnected receiver asks
    for the three codewords of a CONNECT), escalation to the announced codeword count, decodeMCDPSKFrame."""
    from oracle.bindings import McdpskConfig, ZcConfig
    from ria_b200 import mcdpsk, stream
    rng = np.random.default_rng(50 + 10 * connected + spread)
    h = ref.stream_decoder()
    ref.stream_setup_mcdpsk(h, connected, 10, modulation, rate, spread)
    bits = 1 if modulation == DBPSK else 2
    cfg = McdpskConfig.make(bits, (1, 2, 4)[spread], 10)
    rcfg = mcdpsk.MultiCarrierDPSKConfig.from_buffer_copy(bytes(cfg))
    stepper = stream.McdpskStep(rcfg, connected, rate, ctx)
    assert stepper.samples_for_cw(1) == ref.stream_min_control_samples(h)
    pre = len(ref.zc_preamble(ZcConfig.default(), 2)) if connected else 57600
    # "+16" / "-24": the sync position handed to the step is off by that much (a disconnected receiver recovers through
    # its neighbouring-offset retries, :1692-1795)
    kinds = (["ping", "ack", "data3", "data5", "ping", "data3+16", "noise", "ack", "data5-24", "data3"] if not connected
             else ["ack", "data3", "data5", "noise", "ack", "data5", "data3", "data3"])
    txs = []
    for i, kind in enumerate(kinds):
        if kind == "ping":
            tx = ref.stream_encode(2, modulation, rate, 2, b"", 10, spread)
        elif kind == "noise":
            tx = np.zeros(pre + 100000, np.float32)
        else:
            frame = (ref.make_ack_frame("K1ABC", "W2XYZ", i) if kind == "ack" else
                     ref.make_data_frame("K1ABC", "W2XYZ", i, rng.integers(0, 256, size=(30 if kind.startswith("data3") else 70), dtype=np.uint8)))
            tx = ref.stream_encode(2, modulation, rate, 1 if connected else 0, frame, 10, spread)
        txs.append(tx)
    Lw = max(len(t) for t in txs) + 20000
    wins = []
    for kind, tx in zip(kinds, txs):
        p = float(np.mean(tx.astype(np.float64) ** 2)) if kind != "noise" else 0.01
        w = np.zeros(Lw, np.float32)
        w[: len(tx)] = tx
        wins.append(_noisy(w, 12.0, rng, p))
    n = len(kinds)
    x = torch.from_numpy(np.stack(wins)).cuda()
    sync = np.array([pre + (int(k[5:]) if len(k) > 5 and k.startswith("data") else 0) for k in kinds], np.int64)
    cfo = np.zeros(n, np.float32)
    pending = np.zeros(n, np.int32)
    active = np.ones(n, bool)
    seen = dict(ping=0, escalated=0, decoded=0, recovered=0)
    for it in range(4):
        idx = np.nonzero(active)[0]
        if len(idx) == 0:
            break
        got = stepper.step(x[torch.from_numpy(idx).cuda()], sync[idx], cfo[idx], pending[idx])
        for j, i in enumerate(idx):
            res, data = ref.stream_step(h, wins[i], int(sync[i]), 0.0, 10.0, int(pending[i]), 0.0)
            tag = (it, i, kinds[i])
            assert got["state"][j] == res.state, (tag, got["state"][j], res.state, got["pending_total_cw"][j], res.pending_total_cw)
            if res.state == 1:
                assert got["pending_total_cw"][j] == res.pending_total_cw, (tag, got["pending_total_cw"][j], res.pending_total_cw)
                seen["escalated"] += 1
            assert bool(got["has_frame"][j]) == bool(res.has_frame), tag
            if res.has_frame:
                f = res.frame
                assert (got["success"][j], got["is_ping"][j]) == (f.success, f.is_ping), tag
                if f.is_ping:
                    seen["ping"] += 1
                else:
                    assert (got["codewords_ok"][j], got["codewords_failed"][j]) == (f.codewords_ok, f.codewords_failed), tag
                if f.success and not f.is_ping:
                    assert got["frame_type"][j] == f.frame_type and got["frame_len"][j] == f.n_bytes, tag
                    assert bytes(got["frame"][j, : f.n_bytes]) == data, tag
                    seen["decoded"] += 1
                    assert got["sync_pos"][j] == res.sync_pos, (tag, got["sync_pos"][j], res.sync_pos)
                    seen["recovered"] += int(res.sync_pos != sync[i])
            pending[i] = res.pending_total_cw
            active[i] = res.state == 1
    ref.stream_decoder_free(h)
    assert seen["decoded"] >= 6 and seen["escalated"] >= 4, seen
    if not connected:
        assert seen["ping"] == 2, seen

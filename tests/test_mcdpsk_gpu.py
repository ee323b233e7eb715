"""GPU: MC-DPSK batched demod (CFO correction, despreading, two-pass LLRs) vs the unmodified
reference on identical received buffers."""
import numpy as np
import pytest

from oracle.bindings import McdpskConfig, R1_4
from tests.ofdm_common import apply_cfo, awgn, llr_close

pytestmark = pytest.mark.gpu

# (name, bits_per_symbol, spreading, carriers, snr_db)
CASES = [
    ("c3_dbpsk_4x", 1, 4, 10, -8.0),      # BASELINE configs[2]
    ("dbpsk_2x", 1, 2, 10, -4.0),
    ("dbpsk_none", 1, 1, 10, 0.0),
    ("dqpsk_none_8c", 2, 1, 8, 8.0),
    ("dqpsk_2x", 2, 2, 10, 4.0),
]


def _frames(ref, cfg, n, snr_db, rng, n_cw=1, cfo=None):
    out = []
    for i in range(n):
        coded = rng.integers(0, 256, size=81 * n_cw, dtype=np.uint8)     # 648 coded bits per codeword
        tx = ref.mcdpsk_tx_frame(cfg, coded)
        if cfo is not None:
            tx = apply_cfo(tx, cfo[i])
        out.append(awgn(tx, snr_db, rng))
    return out


def _gpu(ctx, cfg, frames, cfo=None, phase=None):
    import torch
    from ria_b200 import mcdpsk
    rcfg = mcdpsk.MultiCarrierDPSKConfig.from_buffer_copy(bytes(cfg))
    dem = mcdpsk.MCDPSKDemodulator(rcfg, ctx)
    x = torch.from_numpy(np.stack(frames)).cuda()
    c = torch.from_numpy(np.asarray(cfo, np.float32)).cuda() if cfo is not None else None
    p = torch.from_numpy(np.asarray(phase, np.float32)).cuda() if phase is not None else None
    out = dem.process_batch(x, c, p)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_matches_reference_no_cfo(ctx, ref, case):
    name, bits, spread, carriers, snr_db = case
    cfg = McdpskConfig.make(bits, spread, carriers)
    rng = np.random.default_rng(sum(map(ord, name)))
    frames = _frames(ref, cfg, 6, snr_db, rng)
    out = _gpu(ctx, cfg, frames)
    for i, rx in enumerate(frames):
        r = ref.mcdpsk_process(cfg, rx, 0.0, 0.0)
        assert r["ready"]
        n = int(out["n_llr"][i])
        assert n == len(r["soft"]) and n >= 648
        got = out["llr"][i, :n]
        ok = llr_close(got, r["soft"])
        assert ok.all(), (i, np.abs(got - r["soft"]).max())
        # every operation of the path is IEEE-exact and atan2f / sinf / cosf are glibc's algorithms
        # restated (csrc/rn_math.h): the soft bits are identical to the reference's
        assert np.array_equal(got.view(np.uint32), r["soft"].view(np.uint32))
        assert abs(out["fading"][i] - r["fading"]) < 1e-5
        assert out["cfo"][i] == r["cfo"]


@pytest.mark.parametrize("case", CASES[:2] + CASES[3:4], ids=[c[0] for c in CASES[:2] + CASES[3:4]])
def test_matches_reference_with_cfo(ctx, ref, case):
    name, bits, spread, carriers, snr_db = case
    cfg = McdpskConfig.make(bits, spread, carriers)
    rng = np.random.default_rng(5 + sum(map(ord, name)))
    cfos = np.array([2.0, -7.5, 19.0, 0.3, -0.05, 33.0], np.float32)
    phases = np.array([0.0, 1.0, -3.0, 3.1, 0.5, -0.2], np.float32)
    frames = _frames(ref, cfg, 6, snr_db + 6, rng, cfo=cfos)
    out = _gpu(ctx, cfg, frames, cfos, phases)
    for i, rx in enumerate(frames):
        r = ref.mcdpsk_process(cfg, rx, float(cfos[i]), float(phases[i]))
        n = int(out["n_llr"][i])
        assert n == len(r["soft"])
        got = out["llr"][i, :n]
        ok = llr_close(got, r["soft"])
        assert ok.all(), (i, cfos[i], np.abs(got - r["soft"]).max())
        # also with the Hilbert CFO correction (fp32 FIR in tap order, exact phase accumulator,
        # glibc's sinf / cosf restated) the soft bits are identical to the reference's
        assert np.array_equal(got.view(np.uint32), r["soft"].view(np.uint32)), (i, cfos[i])
        assert abs(out["fading"][i] - r["fading"]) < 1e-4
        assert out["cfo"][i] == r["cfo"]


def test_multi_codeword_and_edges(ctx, ref):
    import torch
    from ria_b200 import mcdpsk
    cfg = McdpskConfig.make(1, 2, 10)
    rng = np.random.default_rng(1)
    frames = _frames(ref, cfg, 2, 3.0, rng, n_cw=3)
    out = _gpu(ctx, cfg, frames)
    for i, rx in enumerate(frames):
        r = ref.mcdpsk_process(cfg, rx)
        assert int(out["n_llr"][i]) == len(r["soft"]) == 1950
        assert llr_close(out["llr"][i, :1950], r["soft"]).all()
    # ragged lengths: trailing partial symbol / partial spreading group / too short
    rx = frames[0]
    rcfg = mcdpsk.MultiCarrierDPSKConfig.from_buffer_copy(bytes(cfg))
    dem = mcdpsk.MCDPSKDemodulator(rcfg, ctx)
    for n in (4608, 4609, 4608 + 512, 4608 + 3 * 512 + 17, 4000, 4608 + 40 * 512 + 100):
        o = dem.process_batch(torch.from_numpy(rx[:n]).cuda().unsqueeze(0))
        torch.cuda.synchronize()
        r = ref.mcdpsk_process(cfg, rx[:n])
        assert int(o["n_llr"][0]) == len(r["soft"]), n
        if len(r["soft"]):
            assert llr_close(o["llr"][0, : len(r["soft"])].cpu().numpy(), r["soft"]).all(), n
    with pytest.raises(Exception):
        bad = mcdpsk.MultiCarrierDPSKConfig.default()
        bad.samples_per_symbol = 256
        mcdpsk.MCDPSKDemodulator(bad, ctx)


def test_c3_chain_with_ldpc(ctx, ref, port):
    """BASELINE configs[2] shape: DBPSK 10 carriers, 4x spreading, -8 dB: demod -> LDPC R1/4 with
    hard-decision payload bits, success flag and iteration count identical to the reference."""
    import torch
    from ria_b200 import fec
    cfg = McdpskConfig.make(1, 4, 10)
    rng = np.random.default_rng(8)
    frames, sent = [], []
    for i in range(12):
        data = rng.integers(0, 256, size=20, dtype=np.uint8)
        cw = port.ldpc_encode(R1_4, data)[:81]
        frames.append(awgn(ref.mcdpsk_tx_frame(cfg, cw), -8.0, rng))
        sent.append(data)
    out = _gpu(ctx, cfg, frames)
    dec = fec.LDPCDecoder(R1_4, ctx)
    dec.setMaxIterations(50)
    dec.setMinSumFactor(0.9375)
    llr = torch.from_numpy(out["llr"][:, :648].copy()).cuda()
    info, ok, iters = dec.decode_batch(llr)
    torch.cuda.synchronize()
    n_ok = 0
    for i, rx in enumerate(frames):
        r = ref.mcdpsk_process(cfg, rx)
        w_info, w_ok, w_it = ref.ldpc_decode_batch(R1_4, r["soft"][:648], 50, 0.9375, 21)
        assert ok[i].item() == w_ok[0] and iters[i].item() == w_it[0]
        assert np.array_equal(info[i].cpu().numpy(), w_info[0])
        n_ok += int(w_ok[0] and bytes(w_info[0][:20]) == sent[i].tobytes())
    assert n_ok >= 10        # README.md:343: 4x spreading is verified at -8 dB


def test_chirp_acquired_chain_matches_reference(ctx, ref, port):
    """ria_mcdpsk_rx_frames_dev / _host (chirp sync -> process at the detected offset with the
    detected CFO -> chase combining -> LDPC R1/4) against the reference run stage by stage on the
    same rows: training start exact, payload bits / success flag / iteration count exact for the
    first reception and for the chase-combined second reception."""
    import torch
    from ria_b200 import mcdpsk, txsynth
    from ria_b200.sync import SYNC_RESULT_DTYPE
    cfg = McdpskConfig.make(1, 4, 10)
    rcfg = mcdpsk.MultiCarrierDPSKConfig.from_buffer_copy(bytes(cfg))
    rng = np.random.default_rng(21)
    pre = txsynth.chirp_preamble()
    frame_len = None
    rows1, rows2, sent = [], [], []
    lead = 1500
    for i in range(10):
        data = rng.integers(0, 256, size=20, dtype=np.uint8)
        cw = port.ldpc_encode(R1_4, data)[:81]
        body = txsynth.mcdpsk_modulate_frame(rcfg, cw.tobytes())
        frame_len = len(body)
        cfo = float(rng.uniform(-20, 20)) if i % 2 else 0.0
        tx = np.concatenate([np.zeros(lead + 37 * i, np.float32), pre, body, np.zeros(1200 - 37 * i, np.float32)])
        tx = apply_cfo(tx, cfo) if cfo else tx
        rows1.append(awgn(tx, -9.0, rng))
        rows2.append(awgn(tx, -9.0, rng))
        sent.append(data)
    row_len = len(rows1[0])
    window = 120000
    chain = mcdpsk.McdpskRxChain(rcfg, R1_4, 50, 0.9375, 0.15, ctx)
    r1 = chain.process_batch(torch.from_numpy(np.stack(rows1)).cuda(), frame_len, window, None, True)
    torch.cuda.synchronize()
    first = {k: v.cpu().numpy().copy() for k, v in r1.items()}
    r2 = chain.process_batch(torch.from_numpy(np.stack(rows2)).cuda(), frame_len, window, r1["acc"], False)
    torch.cuda.synchronize()
    second = {k: v.cpu().numpy().copy() for k, v in r2.items()}
    host = chain.process_batch_host(np.stack(rows1), frame_len, window)

    after_down = 24000 + 4800
    combined_better = 0
    for i in range(len(rows1)):
        soft = []
        for rows, got in ((rows1, first), (rows2, second)):
            rx = rows[i]
            s = ref.chirp_detect_dual(rx[:window], 0.15)
            g = got["sync"].view(SYNC_RESULT_DTYPE)[i, 0] if got["sync"].ndim == 2 else got["sync"][i]
            assert bool(g["detected"]) == bool(s.detected)
            assert s.detected, "test frames are meant to be detectable"
            assert int(g["aux"]) == int(s.aux) and int(g["start_sample"]) == int(s.start_sample)
            assert abs(float(g["cfo_hz"]) - float(s.cfo_hz)) <= 1e-3 + 1e-4 * abs(s.cfo_hz)
            start = int(s.aux) + after_down
            # the reference demodulator is driven with the CFO the device reported, so that the two
            # chains see identical inputs from here on
            r = ref.mcdpsk_process(cfg, rx[start:start + frame_len], float(g["cfo_hz"]))
            soft.append(r["soft"][:648].astype(np.float32))
        w1 = ref.ldpc_decode_batch(R1_4, soft[0], 50, 0.9375, 24)
        w2 = ref.ldpc_decode_batch(R1_4, (soft[0] + soft[1]).astype(np.float32), 50, 0.9375, 24)
        for w, got in ((w1, first), (w2, second)):
            # soft bits agree within 1e-4, so success / iterations match unless a decode sits on the edge
            assert got["ok"][i] == w[1][0] and got["iters"][i] == w[2][0], (i, got["ok"][i], w[1][0], got["iters"][i], w[2][0])
            assert np.array_equal(got["info"][i, :21], w[0][0][:21])
        assert host["ok"][i] == first["ok"][i] and host["iters"][i] == first["iters"][i]
        assert np.array_equal(host["info"][i], first["info"][i])
        combined_better += int(second["ok"][i]) - int(first["ok"][i])
        if second["ok"][i]:
            assert bytes(second["info"][i, :20]) == sent[i].tobytes()
    assert combined_better >= 0 and second["ok"].sum() >= 8


def test_zc_acquired_retransmission_chain_matches_reference(ctx, ref, port):
    """ria_mcdpsk_zc_rx_frames_dev: a frame first received behind the dual chirp, then retransmitted in connected mode
    behind the Zadoff-Chu data preamble (MCDPSKWaveform::detectDataSync, roots DATA | CONTROL, 31 120-sample window):
    ZC sync result exact, the chase-combined decode equal to the reference run stage by stage on the same rows."""
    import torch
    from oracle.bindings import ZcConfig
    from ria_b200 import mcdpsk, txsynth
    from ria_b200.sync import SYNC_RESULT_DTYPE
    cfg = McdpskConfig.make(1, 4, 10)
    rcfg = mcdpsk.MultiCarrierDPSKConfig.from_buffer_copy(bytes(cfg))
    rng = np.random.default_rng(33)
    chirp = txsynth.chirp_preamble()
    zc = ZcConfig.default()
    zpre = ref.zc_preamble(zc, 2)                                       # DATA root
    rows1, rows2, sent = [], [], []
    frame_len = None
    for i in range(10):
        data = rng.integers(0, 256, size=20, dtype=np.uint8)
        cw = port.ldpc_encode(R1_4, data)[:81]
        body = txsynth.mcdpsk_modulate_frame(rcfg, cw.tobytes())
        frame_len = len(body)
        tx1 = np.concatenate([np.zeros(1500 + 37 * i, np.float32), chirp, body, np.zeros(1200 - 37 * i, np.float32)])
        lead2 = 900 + 211 * i
        tx2 = np.concatenate([np.zeros(lead2, np.float32), zpre, body, np.zeros(4000 - lead2, np.float32)])
        rows1.append(awgn(tx1, -8.0, rng))
        rows2.append(awgn(tx2, -5.0, rng))
        sent.append(data)
    chain = mcdpsk.McdpskRxChain(rcfg, R1_4, 50, 0.9375, 0.15, ctx)
    r1 = chain.process_batch(torch.from_numpy(np.stack(rows1)).cuda(), frame_len, 120000, None, True)
    torch.cuda.synchronize()
    first = {k: v.cpu().numpy().copy() for k, v in r1.items()}
    r2 = chain.process_batch_zc(torch.from_numpy(np.stack(rows2)).cuda(), frame_len, 31120, r1["acc"], False, None, None, 0.2)
    torch.cuda.synchronize()
    second = {k: v.cpu().numpy().copy() for k, v in r2.items()}
    n_ok = n_det = 0
    for i in range(10):
        g1 = first["sync"].view(SYNC_RESULT_DTYPE)[i, 0]
        s1 = ref.chirp_detect_dual(rows1[i][:120000], 0.15)
        assert s1.detected and int(g1["aux"]) == int(s1.aux)
        a = ref.mcdpsk_process(cfg, rows1[i][int(s1.aux) + 28800:int(s1.aux) + 28800 + frame_len], float(g1["cfo_hz"]))["soft"][:648]
        g2 = second["sync"].view(SYNC_RESULT_DTYPE)[i, 0]
        s2 = ref.zc_detect(zc, rows2[i][:31120], 0.2, 4 | 8, 0.0)
        assert bool(g2["detected"]) == bool(s2.detected), i
        if not s2.detected:
            continue
        n_det += 1
        assert int(g2["start_sample"]) == int(s2.start_sample) and np.float32(g2["cfo_hz"]) == np.float32(s2.cfo_hz), i
        st = int(s2.start_sample)
        b = ref.mcdpsk_process(cfg, rows2[i][st:st + frame_len], float(s2.cfo_hz))["soft"][:648]
        w = ref.ldpc_decode_batch(R1_4, (a.astype(np.float32) + b.astype(np.float32)).astype(np.float32), 50, 0.9375, 24)
        assert second["ok"][i] == w[1][0] and second["iters"][i] == w[2][0], (i, second["ok"][i], w[1][0])
        assert np.array_equal(second["info"][i, :21], w[0][0][:21])
        if second["ok"][i]:
            assert bytes(second["info"][i, :20]) == sent[i].tobytes()
            n_ok += 1
    assert n_det >= 6 and n_ok >= 5, (n_det, n_ok)

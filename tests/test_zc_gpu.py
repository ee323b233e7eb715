"""GPU: batched Zadoff-Chu detection vs sync::ZCSync::detect of the unmodified reference.
Mirrors tools/test_zc_sync.cpp: seeded AWGN over -15..20 dB, random offsets, CFO, root
discrimination.  Peak indices / roots / detection flags must be identical."""
import numpy as np
import pytest

from oracle.bindings import ZcConfig
from tests.ofdm_common import apply_cfo

pytestmark = pytest.mark.gpu


def _windows(ref, n, window, snr_dbs, rng, cfo_max=0.0, empty_frac=0.1):
    z = ZcConfig.default()
    pre = {t: ref.zc_preamble(z, t) for t in (0, 1, 2, 3)}
    wins, meta = [], []
    for i in range(n):
        w = np.zeros(window, np.float32)
        t = int(rng.integers(0, 4))
        pos = int(rng.integers(0, window - 2512 - 600))
        snr = float(snr_dbs[i % len(snr_dbs)])
        if rng.random() >= empty_frac:
            sig = pre[t]
            if cfo_max:
                sig = apply_cfo(sig, float(rng.uniform(-cfo_max, cfo_max)))
            w[pos:pos + len(sig)] += sig
            # DPSK-like payload after the preamble
            tail = window - pos - len(sig)
            w[pos + len(sig):] += 0.3 * np.sin(2 * np.pi * 1200 * np.arange(tail) / 48000).astype(np.float32)
        p_sig = 0.8 ** 2 / 2
        w += rng.standard_normal(window).astype(np.float32) * np.float32(np.sqrt(p_sig / 10 ** (snr / 10)))
        wins.append(w)
        meta.append((t, pos))
    return z, wins, meta


def _compare(ref, z, wins, got, threshold, mask, cfos=None):
    n_det = 0
    for i, w in enumerate(wins):
        r = ref.zc_detect(z, w, threshold, mask, float(cfos[i]) if cfos is not None else 0.0)
        g = got[i]
        assert g["detected"] == r.detected, i
        assert g["root"] == r.root and g["frame_type"] == r.frame_type, i
        assert g["start_sample"] == r.start_sample, (i, g["start_sample"], r.start_sample)
        assert abs(g["correlation"] - r.correlation) <= 1e-5 * max(1.0, abs(r.correlation)), i
        if r.detected:
            assert abs(g["cfo_hz"] - r.cfo_hz) <= 1e-3 * max(1.0, abs(r.cfo_hz)), (i, g["cfo_hz"], r.cfo_hz)
            assert abs(g["snr_estimate"] - r.snr_estimate) <= 1e-3
        n_det += r.detected
    return n_det


@pytest.mark.parametrize("window", (7012, 3600))
def test_matches_reference_snr_sweep(ctx, ref, window):
    import torch
    from ria_b200 import sync
    rng = np.random.default_rng(42 + window)
    z, wins, meta = _windows(ref, 40, window, np.arange(-15, 21, 5), rng)
    zs = sync.ZCSync(sync.ZCConfig.from_buffer_copy(bytes(z)), ctx)
    for mask, thr in ((0xF, 0.3), (0xC, 0.25)):
        out = zs.detect_batch(torch.from_numpy(np.stack(wins)).cuda(), thr, mask)
        torch.cuda.synchronize()
        n_det = _compare(ref, z, wins, sync.results(out), thr, mask)
        assert n_det >= 4


def test_matches_reference_with_cfo_and_known_cfo(ctx, ref):
    import torch
    from ria_b200 import sync
    rng = np.random.default_rng(7)
    z, wins, meta = _windows(ref, 24, 7012, [0, 5, 10, 20], rng, cfo_max=20.0, empty_frac=0.0)
    cfos = rng.uniform(-15, 15, size=len(wins)).astype(np.float32)
    cfos[::3] = 0.0
    zs = sync.ZCSync(sync.ZCConfig.from_buffer_copy(bytes(z)), ctx)
    out = zs.detect_batch(torch.from_numpy(np.stack(wins)).cuda(), 0.3, 0xF, torch.from_numpy(cfos).cuda())
    torch.cuda.synchronize()
    assert _compare(ref, z, wins, sync.results(out), 0.3, 0xF, cfos) >= 12


def test_production_window_and_edges(ctx, ref):
    """StreamingDecoder's connected-mode window (31 120 samples, streaming_decoder.cpp:423-435),
    windows shorter than one repetition, empty batch."""
    import torch
    from ria_b200 import sync
    rng = np.random.default_rng(3)
    z, wins, meta = _windows(ref, 6, 31120, [-8, 0, 10], rng, empty_frac=0.0)
    zs = sync.ZCSync(sync.ZCConfig.from_buffer_copy(bytes(z)), ctx)
    out = zs.detect_batch(torch.from_numpy(np.stack(wins)).cuda(), 0.3, 0xC)
    torch.cuda.synchronize()
    _compare(ref, z, wins, sync.results(out), 0.3, 0xC)
    short = torch.zeros((3, 1000), device="cuda")
    res = sync.results(zs.detect_batch(short))
    assert (res["detected"] == 0).all() and (res["start_sample"] == -1).all() and (res["root"] == -1).all()
    assert zs.detect_batch(torch.zeros((0, 7012), device="cuda")).shape[0] == 0

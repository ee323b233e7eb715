"""GPU: OFDM light-preamble (LTS) data sync vs OFDMChirpWaveform::detectDataSync of the reference."""
import numpy as np
import pytest

from oracle.bindings import R1_2, R3_4, DQPSK, QAM64
from tests.ofdm_common import apply_cfo, awgn, make_cfg, tx_frame

pytestmark = pytest.mark.gpu


def _windows(ref, rng, n, window, lead_noise=True, negate_first=False):
    wins, cfos = [], []
    for i in range(n):
        mod, spacing, rate = ((QAM64, 4, R3_4), (DQPSK, 10, R1_2))[i % 2]
        cfg = make_cfg(mod, spacing, 1)
        tx, _, _ = tx_frame(ref, cfg, rate, rng, seq=i)
        if negate_first and i % 3 == 0:
            tx = tx.copy(); tx[:1120] = -tx[:1120]               # burst-interleave marker
        cfo = float(rng.uniform(-8, 8)) if i % 4 else 0.0
        if cfo:
            tx = apply_cfo(tx, cfo)
        w = np.zeros(window, np.float32)
        lead = int(rng.integers(200, 2500)) if lead_noise else 0
        seg = tx[: window - lead]
        w[lead:lead + len(seg)] = seg
        snr = (8, 15, 25)[i % 3]
        p = float(np.mean(tx.astype(np.float64) ** 2))
        w += rng.standard_normal(window).astype(np.float32) * np.float32(np.sqrt(p / 10 ** (snr / 10))) * (0.2 if lead_noise else 1.0)
        wins.append(w)
        cfos.append(np.float32(cfo + rng.uniform(-0.3, 0.3)))
    return wins, np.array(cfos, np.float32)


@pytest.mark.parametrize("lead_noise,negate", [(True, False), (False, False), (True, True)])
def test_matches_reference(ctx, ref, lead_noise, negate):
    import torch
    from ria_b200 import ofdm, sync
    rng = np.random.default_rng(21 + 2 * lead_noise + negate)
    window = 9 * 1120
    wins, cfos = _windows(ref, rng, 18, window, lead_noise, negate)
    rcfg = make_cfg(DQPSK, 10, 1)
    cfg = ofdm.ModemConfig.from_buffer_copy(bytes(rcfg))
    for thr in (0.3, 0.8):
        out = sync.results(sync.ofdm_data_sync_batch(cfg, torch.from_numpy(np.stack(wins)).cuda(),
                                                     torch.from_numpy(cfos).cuda(), thr, ctx))
        n_det = 0
        for i, w in enumerate(wins):
            r = ref.ofdm_data_sync(rcfg, w, float(cfos[i]), thr)
            g = out[i]
            assert g["detected"] == r.detected, (i, g, r.detected, r.correlation)
            assert abs(g["correlation"] - r.correlation) <= 1e-5 * max(1.0, r.correlation), i
            assert g["cfo_hz"] == np.float32(r.cfo_hz)
            if r.detected:
                assert g["start_sample"] == r.start_sample, (i, g["start_sample"], r.start_sample)
                assert g["aux"] == r.aux, i
                n_det += 1
        assert n_det >= 6


def test_edges(ctx, ref):
    import torch
    from ria_b200 import ofdm, sync
    rcfg = make_cfg(DQPSK, 10, 1)
    cfg = ofdm.ModemConfig.from_buffer_copy(bytes(rcfg))
    short = sync.results(sync.ofdm_data_sync_batch(cfg, torch.zeros((2, 3000), device="cuda"), None, 0.3, ctx))
    assert (short["detected"] == 0).all() and (short["start_sample"] == -1).all()
    z = sync.results(sync.ofdm_data_sync_batch(cfg, torch.zeros((1, 6000), device="cuda"), None, 0.3, ctx))
    r = ref.ofdm_data_sync(rcfg, np.zeros(6000, np.float32), 0.0, 0.3)
    assert z["detected"][0] == r.detected == 0 and z["correlation"][0] == r.correlation

"""CPU: the plain-C restatement of the Schmidl-Cox timing metric (oracle/ofdm_sync_oracle.c) against the unmodified
reference (oracle/_ref, Impl::measureCorrelation) and against the committed golden vectors -- bit for bit."""
import os

import numpy as np
import pytest

from oracle.bindings import DQPSK, QAM64, Port
from tests.ofdm_common import make_cfg

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cox_golden.npz")
CP = 96          # ModemConfig::getCyclicPrefix() of the configurations in the fixture (CP mode MEDIUM, FFT 1024)


@pytest.fixture(scope="module")
def port():
    return Port()


@pytest.mark.parametrize("name", ["qam64_sp4", "dqpsk_sp5"])
def test_port_metric_matches_golden(port, name):
    g = np.load(GOLD)
    x = g[f"{name}_win"].astype(np.float32)
    for i in range(len(x)):
        m, _, _, _ = port.cox_correlation(x[i], int(g[f"{name}_corr_off"][i]), CP)
        assert m.view(np.uint32) == g[f"{name}_corr"][i].view(np.uint32), (name, i, m, g[f"{name}_corr"][i])


@pytest.mark.parametrize("name,mod,spacing", [("qam64_sp4", QAM64, 4), ("dqpsk_sp5", DQPSK, 5)])
def test_port_metric_matches_reference(port, ref, name, mod, spacing):
    g = np.load(GOLD)
    cfg = make_cfg(mod, spacing, 1)
    x = g[f"{name}_win"].astype(np.float32)
    rng = np.random.default_rng(5)
    n = 0
    for i in range(len(x)):
        for off in [0, 1, 63] + rng.integers(0, len(x[i]) - 1200, size=12).tolist() + [len(x[i]) - 1120, len(x[i]) - 1119, len(x[i]) - 10]:
            want = np.float32(ref.ofdm_cox_correlation(cfg, x[i], int(off)))
            got, _, _, _ = port.cox_correlation(x[i], int(off), CP)
            assert got.view(np.uint32) == want.view(np.uint32), (name, i, off, got, want)
            n += 1
    assert n >= 100


def test_port_metric_edges(port):
    z = np.zeros(4096, np.float32)
    assert port.cox_correlation(z, 0, CP)[0] == 0.0                      # silence: normalisation below 1e-10
    assert port.cox_correlation(z[:1000], 0, CP)[0] == 0.0               # window does not fit
    dc = np.full(4096, 0.25, np.float32)
    assert port.cox_correlation(dc, 100, CP)[0] == 0.0                   # pure DC is removed before the transform


@pytest.mark.parametrize("name,mod,spacing", [("qam64_sp4", QAM64, 4), ("dqpsk_sp5", DQPSK, 5)])
def test_port_coarse_cfo_matches_reference(port, ref, name, mod, spacing):
    """orc_cox_coarse_cfo against Impl::estimateCoarseCFO (through the checker's refineLTSTiming tap)"""
    g = np.load(GOLD)
    cfg = make_cfg(mod, spacing, 1)
    x = g[f"{name}_win"].astype(np.float32)
    rng = np.random.default_rng(9)
    n = clamped = 0
    for i in range(len(x)):
        for off in rng.integers(0, len(x[i]) - 8000, size=10).tolist() + [len(x[i]) - 1120, len(x[i]) - 1119]:
            _, want = ref.ofdm_cox_refine_lts(cfg, x[i], int(off))
            got = port.cox_coarse_cfo(x[i], int(off), CP)
            assert got.view(np.uint32) == np.float32(want).view(np.uint32), (name, i, off, got, want)
            n += 1
            clamped += abs(float(got)) == 46.0
    assert n >= 90

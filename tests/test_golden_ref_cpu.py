"""CPU: the committed OFDM_COX golden vectors (tests/golden/cox_golden.npz) against the unmodified reference
(oracle/_ref) where it is built -- pins the fixture to the reference it was generated from
(tests/golden/make_golden.py cox), so a stale or hand-edited fixture cannot go unnoticed."""
import os

import numpy as np
import pytest

from oracle.bindings import DQPSK, QAM64
from tests.ofdm_common import make_cfg

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cox_golden.npz")


@pytest.mark.parametrize("name,mod,spacing", [("qam64_sp4", QAM64, 4), ("dqpsk_sp5", DQPSK, 5)])
def test_cox_golden_is_what_the_reference_computes(ref, name, mod, spacing):
    g = np.load(GOLD)
    cfg = make_cfg(mod, spacing, 1)
    x = g[f"{name}_win"].astype(np.float32)
    for tag in ("a", "b"):
        thr, nf_in = float(g[f"{name}_{tag}_thr"]), float(g[f"{name}_{tag}_nf_in"])
        for i in range(len(x)):
            f, pos, cfo, nf = ref.ofdm_cox_search_sync(cfg, x[i], thr, nf_in)
            assert int(f) == int(g[f"{name}_{tag}_found"][i])
            assert np.float32(nf).view(np.uint32) == g[f"{name}_{tag}_nf_out"][i].view(np.uint32)
            if f:
                assert pos == int(g[f"{name}_{tag}_pos"][i])
                assert np.float32(cfo).view(np.uint32) == g[f"{name}_{tag}_cfo"][i].view(np.uint32)
    for i in range(len(x)):
        c = ref.ofdm_cox_correlation(cfg, x[i], int(g[f"{name}_corr_off"][i]))
        assert np.float32(c).view(np.uint32) == g[f"{name}_corr"][i].view(np.uint32)
    assert 1 <= int(g[f"{name}_a_found"].sum()) < len(x)          # the set holds hits and misses

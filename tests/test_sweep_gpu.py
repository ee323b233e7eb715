"""GPU: the adaptive-waveform SNR sweep driver (BASELINE configs[4]) end to end on one GPU:
selection ladder -> host TX synthesis -> on-device Watterson/AWGN channel -> the receive chain of
the selected mode -> counters."""
import argparse

import pytest

pytestmark = pytest.mark.gpu


def test_sweep_awgn_is_error_free_where_the_ladder_says_so():
    import sweep
    args = argparse.Namespace(frames=192, condition="awgn", snr_min=-12.0, snr_max=30.0, snr_step=6.0, quiet=True)
    res = sweep.run_sweep(args)
    assert len(res) == 8
    kinds = {r["mode"].split()[0] for r in res}
    assert kinds == {"MC-DPSK", "OFDM"}                    # both waveform families were exercised
    for r in res:
        assert r["frames"] == 192
        # the ladder's thresholds carry a few dB of margin on a static channel
        assert r["fer"] <= 0.05, r
    assert res[-1]["estimated_throughput_bps"] > res[0]["estimated_throughput_bps"]


def test_sweep_fading_degrades_gracefully():
    import sweep
    args = argparse.Namespace(frames=128, condition="moderate", snr_min=-4.0, snr_max=28.0, snr_step=16.0, quiet=True)
    res = sweep.run_sweep(args)
    assert [r["frames"] for r in res] == [128, 128, 128]
    assert all(0.0 <= r["fer"] <= 1.0 for r in res)

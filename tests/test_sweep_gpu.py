"""GPU: the adaptive-waveform SNR sweep driver (BASELINE configs[4]) end to end on one GPU:
selection ladder -> host TX synthesis -> on-device Watterson/AWGN channel -> the receive chain of
the selected mode -> counters."""
import argparse

import pytest

pytestmark = pytest.mark.gpu


def test_sweep_awgn_is_error_free_where_the_ladder_says_so():
    import sweep
    args = argparse.Namespace(frames=192, condition="awgn", snr_min=-12.0, snr_max=30.0, snr_step=6.0, quiet=True,
                              first_pass_only=True)
    res = sweep.run_sweep(args)
    assert len(res) == 8
    # The complete decodeFixedFrame (default) drops every frame one of whose codeword chunks 1..3 starts
    # with 0xD5: CodewordStatus::reassemble (frame_v2.cpp:974) takes it for a DATA_CW_MARKER, the frame no
    # longer verifies, the repair fails and all four codewords are marked failed (:1879-1884).  That is the
    # reference's behaviour (tests/test_ldpc_retry_gpu.py pins it), 3/256 of random frames; with a pool of 16
    # distinct frames per mode it costs whole multiples of 1/16.
    args_full = argparse.Namespace(**{**vars(args), "first_pass_only": False})
    full = sweep.run_sweep(args_full)
    for a, b in zip(res, full):
        assert b["frames_ok"] <= a["frames_ok"] + 2 and b["fer"] <= a["fer"] + 2.0 / 16 + 0.02, (a, b)
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

"""CPU: the waveform / rate selection ladder against the reference on a dense (SNR, fading) grid."""
import numpy as np


def test_recommend_matches_reference(ria_lib, ref):
    from ria_b200 import selection
    for snr in np.arange(-16.0, 32.0, 0.5):
        for fading in (0.0, 0.04, 0.05, 0.1, 0.149, 0.15, 0.3, 0.45, 0.649, 0.65, 0.9, 1.09, 1.1, 1.5):
            a = selection.recommendWaveformAndRate(snr, fading)
            b = ref.recommend_waveform(float(snr), float(fading))
            for f in ("waveform", "modulation", "rate", "estimated_throughput_bps", "num_carriers", "spreading"):
                assert getattr(a, f) == getattr(b, f), (snr, fading, f)
            for wf in (selection.MC_DPSK, selection.OFDM_CHIRP, selection.OFDM_COX):
                a = selection.recommendDataMode(snr, wf, fading)
                b = ref.recommend_data_mode(float(snr), wf, float(fading))
                for f in ("modulation", "rate", "num_carriers", "spreading"):
                    assert getattr(a, f) == getattr(b, f), (snr, fading, wf, f)

"""protocol::recommendWaveformAndRate / recommendDataMode (src/protocol/waveform_selection.hpp)
through the library's host-side mirror."""
from __future__ import annotations

import ctypes as C

from ._lib import lib

OFDM_COX, OTFS_EQ, OTFS_RAW, MFSK, MC_DPSK, OFDM_CHIRP = range(6)


class WaveformRecommendation(C.Structure):
    _fields_ = [("waveform", C.c_int32), ("modulation", C.c_int32), ("rate", C.c_int32),
                ("estimated_throughput_bps", C.c_float), ("num_carriers", C.c_int32), ("spreading", C.c_int32)]


def recommendWaveformAndRate(snr_db: float, fading_index: float) -> WaveformRecommendation:
    r = WaveformRecommendation()
    lib().ria_recommend_waveform(float(snr_db), float(fading_index), C.addressof(r))
    return r


def recommendDataMode(snr_db: float, waveform: int, fading_index: float = 0.0) -> WaveformRecommendation:
    r = WaveformRecommendation()
    lib().ria_recommend_data_mode(float(snr_db), int(waveform), float(fading_index), C.addressof(r))
    return r

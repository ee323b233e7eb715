"""Shared helpers for the OFDM parity tests (test infrastructure)."""
import numpy as np

from oracle.bindings import (ModemConfig, R1_2, R1_4, R2_3, R3_4, BITS_PER_CARRIER, BYTES_PER_CW,
                             DBPSK, BPSK, DQPSK, QPSK, D8PSK, QAM16, QAM32, QAM64, QAM256)

# (name, modulation, pilot_spacing, use_pilots, code rate, snr_db)
CASES = [
    ("c4_qam64_r34_sp4", QAM64, 4, 1, R3_4, 28.0),      # BASELINE configs[3]
    ("c1_dqpsk_r12_sp10", DQPSK, 10, 1, R1_2, 15.0),    # BASELINE configs[0]
    ("qam16_r23_sp5", QAM16, 5, 1, R2_3, 22.0),
    ("qpsk_r12_sp5", QPSK, 5, 1, R1_2, 12.0),
    ("qam32_r34_sp8", QAM32, 8, 1, R3_4, 26.0),
    ("bpsk_r14_sp5", BPSK, 5, 1, R1_4, 6.0),
    ("dbpsk_r14_sp10", DBPSK, 10, 1, R1_4, 6.0),
    ("dqpsk_nopilots", DQPSK, 2, 0, R1_2, 18.0),
    ("d8psk_r12_sp8", D8PSK, 8, 1, R1_2, 20.0),         # demapD8PSK (soft_demap.hpp:238-263), single pass on AWGN
    ("qam256_r34_sp4", QAM256, 4, 1, R3_4, 36.0),       # demapQAM256 (soft_demap.hpp:145-164)
]


def make_cfg(mod, spacing, use_pilots):
    return ModemConfig.make(mod, spacing, use_pilots)


def tx_frame(ref, cfg, rate, rng, seq=1):
    """payload -> v2 data frame -> encodeFixedFrame -> OFDM TX samples (reference TX chain)."""
    bpc = BYTES_PER_CW[rate]
    payload = rng.integers(0, 256, size=4 * bpc - 19 - int(rng.integers(0, 8)), dtype=np.uint8)
    frame = ref.make_data_frame("K1ABC", "W2XYZ", seq, payload)
    bps = cfg.data_carriers() * BITS_PER_CARRIER[cfg.modulation]
    coded = ref.encode_fixed_frame(frame, rate, True, bps)
    return ref.ofdm_tx_frame(cfg, coded), frame, bps


def awgn(tx, snr_db, rng):
    p = float(np.mean(tx.astype(np.float64) ** 2))
    sigma = np.sqrt(p / 10 ** (snr_db / 10))
    return (tx + rng.standard_normal(len(tx)).astype(np.float32) * np.float32(sigma)).astype(np.float32)


def apply_cfo(x, cfo_hz, fs=48000.0):
    """Frequency-shift a real passband signal (analytic signal via FFT Hilbert)."""
    from scipy.signal import hilbert
    n = np.arange(len(x))
    return np.real(hilbert(x.astype(np.float64)) * np.exp(2j * np.pi * cfo_hz * n / fs)).astype(np.float32)


def llr_close(a, b, rtol=1e-4, atol=1e-4):
    """north_star tolerance: soft LLRs within 1e-4 relative in fp32.  LLRs are clipped to
    [-20, 20] and many are a large scale (2/noise_var, several hundred) times a difference of
    nearly equal terms, so a purely relative bound is undefined near zero: the absolute floor is
    1e-4 (5e-6 of the LLR full scale)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) <= atol + rtol * np.abs(b)

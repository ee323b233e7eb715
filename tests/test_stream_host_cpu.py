"""CPU: the pure host helpers of ria_b200/stream.py (the batched StreamingDecoder step) against the reference:
_parse_header vs v2::parseHeader (frame_v2.cpp:1195-1253) on data / control frames and corrupted copies, and
_initial_phase vs the float expression of OFDMChirpWaveform::process (ofdm_chirp_waveform.cpp:402-411)."""
import numpy as np

from ria_b200.stream import _initial_phase, _parse_header


def test_parse_header_matches_reference(ref):
    rng = np.random.default_rng(11)
    rows = []
    for i in range(40):
        f = ref.make_data_frame("K1ABC", "W2XYZ", i, rng.integers(0, 256, size=int(rng.integers(1, 200)), dtype=np.uint8))
        rows.append(np.frombuffer(f, np.uint8)[:20].copy())
    for i in range(12):
        f = ref.make_ack_frame("K1ABC", "W2XYZ", i, nack=bool(i & 1))
        rows.append(np.frombuffer(f, np.uint8)[:20].copy())
    base = len(rows)
    for k in range(base):                                   # one flipped bit somewhere in the 20 bytes
        r = rows[k].copy()
        bit = int(rng.integers(0, 160))
        r[bit // 8] ^= np.uint8(1 << (bit % 8))
        rows.append(r)
    rows.append(np.zeros(20, np.uint8))
    rows.append(rng.integers(0, 256, size=20, dtype=np.uint8))
    d0 = np.stack(rows)
    valid, ftype, total_cw, is_control = _parse_header(d0)
    n_valid = 0
    for i, r in enumerate(d0):
        st = ref.parse_header(r)
        assert bool(valid[i]) == bool(st.header_valid), (i, r[:4])
        if st.header_valid:
            n_valid += 1
            assert int(ftype[i]) == int(st.type) and int(total_cw[i]) == int(st.total_cw), (i, ftype[i], st.type)
    assert n_valid >= base                                   # every untouched frame parses; most corrupted ones do not
    assert n_valid < len(d0) - base // 2


def test_initial_phase_rounding_steps():
    """Regression pin of the float / double rounding steps (the end-to-end check against the reference waveform is the
    GPU harness, tests/test_waveform_dropin_gpu.py): double expression -> float, wrap in double steps -> float."""
    rng = np.random.default_rng(3)
    cfo = rng.uniform(-30, 30, size=200).astype(np.float32)
    pos = rng.integers(0, 200000, size=200).astype(np.int64)
    got = _initial_phase(cfo, pos, 48000.0)
    for i in range(len(cfo)):
        # float initial_phase = -2.0f * M_PI * cfo_hz_ * training_start_sample / sample_rate: evaluated in double, then
        # wrapped with `while (phase > M_PI) phase -= 2.0f * M_PI` on a float variable
        p = np.float32(-2.0 * np.pi * float(cfo[i]) * float(pos[i]) / 48000.0)
        while float(p) > np.pi:
            p = np.float32(float(p) - 2.0 * np.pi)
        while float(p) < -np.pi:
            p = np.float32(float(p) + 2.0 * np.pi)
        assert got[i].view(np.uint32) == p.view(np.uint32)
        assert -np.pi - 1e-6 <= float(got[i]) <= np.pi + 1e-6

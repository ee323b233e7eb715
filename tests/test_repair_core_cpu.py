"""CPU: the serial core of decodeFixedFrame's false-positive repair (ria_b200/csrc/frame_repair_core.h,
the same source the device runs) against libstdc++'s std::sort and against the unmodified reference."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle.bindings import R1_2, R2_3, R3_4, BYTES_PER_CW, awgn_llrs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
MAX_ITER = {R1_2: 80, R2_3: 70, R3_4: 60}


@pytest.fixture(scope="module")
def core(tmp_path_factory):
    so = tmp_path_factory.mktemp("rc") / "librepair_core.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(so),
                    os.path.join(ROOT, "tests", "repair_core_host.cpp")], check=True)
    L = C.CDLL(str(so))
    L.rc_sort.argtypes = [_f32p, _u16p, C.c_int]
    L.rc_std_sort.argtypes = [_f32p, _u16p, C.c_int]
    L.rc_frame_valid.argtypes = [_u8p, C.c_int]
    L.rc_repair_bitflips.argtypes = [_u8p, C.c_int, _f32p]
    return L


def test_sort_is_libstdcxx_sort_including_ties(core):
    rng = np.random.default_rng(5)
    cases = []
    for n in list(range(0, 40)) + [100, 333, 960, 1904, 1920]:
        for levels in (0, 2, 5, 50):
            key = rng.random(n).astype(np.float32) * 20
            if levels:
                key = (np.floor(key / 20 * levels) / levels * 20).astype(np.float32)      # heavy ties
            cases.append(key)
    # patterns that stress the partition: sorted, reversed, organ pipe, constant, clipped soft bits
    for n in (17, 64, 500, 1920):
        a = np.arange(n, dtype=np.float32)
        cases += [a, a[::-1].copy(), np.concatenate([a[: n // 2], a[: n - n // 2][::-1]]), np.zeros(n, np.float32),
                  np.minimum(np.abs(rng.standard_normal(n)).astype(np.float32) * 15, 20.0).astype(np.float32),
                  np.maximum(np.abs(rng.standard_normal(n)).astype(np.float32), 0.01).astype(np.float32) * (rng.random(n) < .5)]
    for key in cases:
        key = np.ascontiguousarray(key, np.float32)
        n = len(key)
        k1, k2 = key.copy(), key.copy()
        v1 = np.arange(n, dtype=np.uint16)
        v2 = v1.copy()
        core.rc_sort(k1, v1, n)
        core.rc_std_sort(k2, v2, n)
        assert np.array_equal(k1, k2)
        assert np.array_equal(v1, v2), n


def _median_of_three_killer(n):
    """Musser's sequence: drives median-of-3 quicksort to its depth limit (heap-sort fallback)."""
    k = n // 2
    a = np.zeros(n, np.float32)
    for i in range(1, k + 1):
        if i % 2 == 1:
            a[i - 1] = i
            a[i] = k + i
        a[k + i - 1] = 2 * i
    return a


def test_sort_heap_fallback_path(core):
    for n in (64, 256, 1024, 1920):
        key = _median_of_three_killer(n)
        k1, k2 = key.copy(), key.copy()
        v1 = np.arange(n, dtype=np.uint16)
        v2 = v1.copy()
        core.rc_sort(k1, v1, n)
        core.rc_std_sort(k2, v2, n)
        assert np.array_equal(k1, k2) and np.array_equal(v1, v2), n


def _deinterleave(soft, rate, bps):
    """FrameInterleaver::deinterleave + ChannelInterleaver::deinterleave (numpy restatement used by
    tests only; the device gather is checked against the reference in tests/test_ofdm_gpu.py)."""
    import ria_b200
    step = ria_b200.lib().ria_channel_interleaver_step(bps, 648)
    b = np.arange(648)
    out = np.empty((4, 648), np.float32)
    for c in range(4):
        cw = soft[4 * b + (c + b) % 4]
        out[c] = cw[(b * step) % 648]
    return out


@pytest.mark.parametrize("rate,esn0,bps", [(R3_4, 7.5, 264), (R2_3, 5.5, 176), (R1_2, 3.4, 106)])
def test_false_positive_repair_matches_reference(core, ref, ria_lib, rate, esn0, bps):
    """Frames whose four codewords all pass parity in the first pass but whose frame is invalid: the
    reference's complete decodeFixedFrame (bit-flip searches, then the re-decode fallback) vs the
    restated core (+ the fallback composed from the reference's own decoder)."""
    rng = np.random.default_rng(900 + rate)
    bpc = BYTES_PER_CW[rate]
    n_checked = n_flip = n_redecode = n_fail = 0
    for i in range(260):
        payload = rng.integers(0, 256, size=4 * bpc - 19 - int(rng.integers(0, 8)), dtype=np.uint8)
        if i % 9 == 0:
            payload[bpc - 17 + int(rng.integers(0, 2)) * bpc] = 0xD5      # DATA_CW_MARKER quirk of reassemble()
        frame = ref.make_data_frame("K1ABC", "W2XYZ", i, payload)
        coded = ref.encode_fixed_frame(frame, rate, True, bps)
        soft = awgn_llrs(np.unpackbits(coded)[:2592], esn0, rng)
        data, ok, _ = ref.frame_decode_first_pass(soft, rate, True, bps)
        if not ok.all():
            continue
        data = np.ascontiguousarray(data)
        if core.rc_frame_valid(data, bpc):
            continue
        cw_soft = np.ascontiguousarray(_deinterleave(soft, rate, bps))
        got = data.copy()
        recovered = bool(core.rc_repair_bitflips(got, bpc, cw_soft))
        n_flip += recovered
        if not recovered:
            # fallback (:1848-1876): re-decode with other min-sum factors, keep a different codeword if the frame verifies
            for factor in (0.75, 0.625, 0.5, 0.875):
                for c in range(4):
                    if recovered:
                        break
                    b, s, _it = ref.ldpc_decode_soft(rate, cw_soft[c], MAX_ITER[rate], factor)
                    if s and not np.array_equal(b[:bpc], got[c * bpc:(c + 1) * bpc]):
                        keep = got[c * bpc:(c + 1) * bpc].copy()
                        got[c * bpc:(c + 1) * bpc] = b[:bpc]
                        if core.rc_frame_valid(got, bpc):
                            recovered = True
                            n_redecode += 1
                        else:
                            got[c * bpc:(c + 1) * bpc] = keep
        w_data, w_ok = ref.decode_fixed_frame_full(soft, rate, True, bps)
        assert bool(w_ok.all()) == recovered, i
        assert not w_ok.any() or w_ok.all()
        if recovered:
            assert np.array_equal(w_data, got), i
        else:
            n_fail += 1
        n_checked += 1
    assert n_checked >= 20, n_checked
    assert n_flip + n_redecode > 0
    print(f"rate {rate}: {n_checked} false-positive frames, {n_flip} repaired by bit flips, "
          f"{n_redecode} by re-decode, {n_fail} given up")

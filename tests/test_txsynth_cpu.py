"""CPU: the numpy TX synthesiser (ria_b200/txsynth.py, used for bench/test inputs) produces
frames the UNMODIFIED reference receiver decodes, and matches the reference TX closely."""
import numpy as np
import pytest

from oracle.bindings import BITS_PER_CARRIER, ModemConfig as RefCfg, R1_2, R3_4, DQPSK, QAM64, QAM16, R2_3


@pytest.mark.parametrize("mod,spacing,rate", [(QAM64, 4, R3_4), (DQPSK, 10, R1_2), (QAM16, 5, R2_3)])
def test_txsynth_decodes_under_reference(ria_lib, ref, mod, spacing, rate):
    from ria_b200 import ofdm, txsynth
    cfg = ofdm.ModemConfig.default(mod, use_pilots=1, pilot_spacing=spacing)
    rcfg = RefCfg.from_buffer_copy(bytes(cfg))
    pool, raw = txsynth.make_frame_pool(cfg, rate, 3, seed=4)
    bps = cfg.getDataCarriers() * BITS_PER_CARRIER[mod]
    rng = np.random.default_rng(0)
    for tx, fr in zip(pool, raw):
        # frame bytes and coded bits agree with the reference TX chain
        assert ref.make_data_frame("K1ABC", "W2XYZ", (fr[4] << 8) | fr[5], fr[17:-2]) == fr
        coded = np.unpackbits(ref.encode_fixed_frame(fr, rate, True, bps))[:2592]
        assert np.array_equal(coded, txsynth.encode_fixed_frame_bits(fr, rate, True, bps))
        ref_tx = ref.ofdm_tx_frame(rcfg, np.packbits(coded))
        assert ref_tx.shape == tx.shape
        assert np.abs(ref_tx - tx).max() < 2e-5 * np.abs(ref_tx).max()
        p = float(np.mean(tx.astype(np.float64) ** 2))
        rx = (tx + rng.standard_normal(len(tx)).astype(np.float32) * np.sqrt(p / 10 ** 2.8)).astype(np.float32)
        r = ref.ofdm_process_presynced(rcfg, rx)
        data, ok, _ = ref.frame_decode_first_pass(r["soft"], rate, True, bps)
        assert ok.all() and bytes(data[: len(fr)]) == fr
        assert ref.parse_header(data).frame_crc_ok == 1


def test_host_preambles_are_the_reference_samples(ref, ria_lib):
    """ria_zc_preamble_host / ria_chirp_generate_host vs ZCSync::generatePreamble and
    ChirpSync::generate of the unmodified reference: bit-identical."""
    from oracle.bindings import ZcConfig
    from ria_b200 import sync
    zc = ZcConfig.default()
    for frame_type, root in ((0, zc.root_ping), (1, zc.root_pong), (2, zc.root_data), (3, zc.root_control)):
        want = ref.zc_preamble(zc, frame_type)
        got = sync.zc_preamble_host(sync.ZCConfig.default(), root)
        assert len(got) == len(want)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), frame_type
    want = ref.chirp_generate()
    got = sync.chirp_generate_host()
    assert len(got) == len(want) and np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_vectorised_frame_builder_matches_the_reference(ref):
    """txsynth.make_data_frames (bench / sweep inputs): byte-identical to DataFrame::makeData + serialize, and
    the 0xD5 chunk-start avoidance keeps the frame valid."""
    from ria_b200 import txsynth
    rng = np.random.default_rng(9)
    payloads = rng.integers(0, 256, size=(40, 219), dtype=np.uint8)
    frames = txsynth.make_data_frames("K1ABC", "W2XYZ", 65520, payloads)
    for i in range(len(payloads)):
        assert bytes(frames[i]) == ref.make_data_frame("K1ABC", "W2XYZ", (65520 + i) & 0xFFFF, payloads[i])
    payloads[:, 60 - 17] = 0xD5
    frames = txsynth.make_data_frames("K1ABC", "W2XYZ", 0, payloads, bytes_per_cw=60)
    assert (frames[:, 60] == 0xD4).all()
    for i in (0, 7, 39):
        st = ref.parse_header(bytes(frames[i]))
        assert st.header_valid and st.frame_crc_ok

"""GPU parity of the on-device transmit synthesis against the unmodified reference:
v2::encodeFixedFrame bytes and OFDMModulator samples must be identical."""
import numpy as np
import pytest
import torch

from tests.ofdm_common import CASES, make_cfg
from oracle.bindings import BITS_PER_CARRIER, BYTES_PER_CW, R1_4, R1_2, R2_3, R3_4

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rate,bps,use_ci", [(R1_4, 53, True), (R1_2, 106, True), (R2_3, 176, True), (R3_4, 264, True), (R3_4, 264, False)])
def test_encode_fixed_frame_is_byte_identical(ctx, ref, rate, bps, use_ci):
    from ria_b200 import ofdm
    rng = np.random.default_rng(rate)
    bpc = BYTES_PER_CW[rate]
    n = 12
    frames = np.zeros((n, 4 * bpc), np.uint8)
    lens = []
    for i in range(n):
        ln = 4 * bpc - int(rng.integers(0, 40)) if i else 4 * bpc
        frames[i, :ln] = rng.integers(0, 256, size=ln, dtype=np.uint8)
        lens.append(ln)
    got = ofdm.encode_fixed_frame_batch(torch.from_numpy(frames).cuda(), rate, use_ci, bps, ctx).cpu().numpy()
    for i in range(n):
        want = ref.encode_fixed_frame(bytes(frames[i, :lens[i]]), rate, use_ci, bps)
        assert np.array_equal(got[i], want), i


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_ofdm_tx_is_sample_identical(ctx, ref, case):
    from ria_b200 import ofdm
    name, mod, spacing, use_pilots, rate, _snr = case
    cfg_o = make_cfg(mod, spacing, use_pilots)
    cfg = ofdm.ModemConfig.from_buffer_copy(bytes(cfg_o))
    rng = np.random.default_rng(77)
    coded = rng.integers(0, 256, size=(6, 324), dtype=np.uint8)
    got = ofdm.ofdm_tx_frames(cfg, torch.from_numpy(coded).cuda(), ctx).cpu().numpy()
    for i in range(len(coded)):
        want = ref.ofdm_tx_frame(cfg_o, coded[i])
        assert got.shape[1] == len(want), (got.shape, len(want))
        assert np.array_equal(got[i].view(np.uint32), want.view(np.uint32)), (name, i, float(np.abs(got[i] - want).max()))


def test_device_tx_chain_decodes_and_matches_host_payload(ctx, ref):
    """frames -> encodeFixedFrame -> OFDM TX on the device -> device receive chain: payload back, CRC valid."""
    from ria_b200 import ofdm
    cfg = ofdm.ModemConfig.high_throughput(ofdm.QAM64)
    rng = np.random.default_rng(5)
    n = 64
    frames = np.stack([np.frombuffer(ref.make_data_frame("K1ABC", "W2XYZ", i, rng.integers(0, 256, size=221, dtype=np.uint8)), np.uint8)
                       for i in range(n)])
    bps = cfg.getDataCarriers() * 6
    coded = ofdm.encode_fixed_frame_batch(torch.from_numpy(frames).cuda(), R3_4, True, bps, ctx)
    tx = ofdm.ofdm_tx_frames(cfg, coded, ctx)
    chain = ofdm.OfdmRxChain(cfg, R3_4, True, ctx)
    data, status, _ = chain.process_batch(tx)
    st = ofdm.status_array(status)
    assert st["all_ok"].all() and st["frame_crc_ok"].all()
    assert np.array_equal(data.cpu().numpy()[:, :frames.shape[1]], frames)


@pytest.mark.parametrize("bits,spreading,carriers", [(1, 4, 10), (1, 2, 10), (1, 1, 10), (2, 1, 10), (2, 1, 5), (1, 4, 8)])
def test_mcdpsk_tx_is_sample_identical(ctx, ref, bits, spreading, carriers):
    """MultiCarrierDPSKModulator training + reference + modulate(data) on the device vs the reference."""
    from oracle.bindings import McdpskConfig
    from ria_b200 import mcdpsk
    cfg_o = McdpskConfig.make(bits, spreading, carriers)
    cfg = mcdpsk.MultiCarrierDPSKConfig.from_buffer_copy(bytes(cfg_o))
    rng = np.random.default_rng(bits * 100 + spreading * 10 + carriers)
    data = rng.integers(0, 256, size=(5, 81), dtype=np.uint8)
    got = mcdpsk.mcdpsk_tx_frames(cfg, torch.from_numpy(data).cuda(), ctx).cpu().numpy()
    for i in range(len(data)):
        want = ref.mcdpsk_tx_frame(cfg_o, data[i])
        assert got.shape[1] == len(want), (got.shape, len(want))
        assert np.array_equal(got[i].view(np.uint32), want.view(np.uint32)), (i, float(np.abs(got[i] - want).max()))


def test_new_entry_points_reject_bad_arguments_and_accept_empty_batches(ctx):
    """Error behaviour of the round-1 additions: negative status + message, never a crash; n = 0 is a no-op."""
    import ctypes as C
    import ria_b200
    from ria_b200 import ofdm, mcdpsk
    L = ria_b200.lib()
    h = ctx.handle
    cfg = ofdm.ModemConfig.high_throughput(ofdm.QAM64)
    buf = torch.zeros(4096, dtype=torch.uint8, device="cuda")
    out = torch.zeros(1 << 16, dtype=torch.float32, device="cuda")
    # empty batches
    assert L.ria_encode_fixed_frame_batch_dev(h, 4, 1, 264, buf.data_ptr(), 240, 240, 0, buf.data_ptr()) == 0
    assert L.ria_ofdm_tx_frames_dev(h, C.addressof(cfg), buf.data_ptr(), 324, 324, 0, out.data_ptr(), 13440) == 0
    assert L.ria_burst_deinterleave_batch_dev(h, out.data_ptr(), 2592, 4, 0, out.data_ptr(), 2592) == 0
    assert L.ria_ldpc_robust_decode_batch_dev(h, 4, out.data_ptr(), 0, buf.data_ptr(), 61, buf.data_ptr(), out.data_ptr(), None) == 0
    # bad arguments -> negative status and a message
    assert L.ria_encode_fixed_frame_batch_dev(h, 99, 1, 264, buf.data_ptr(), 240, 240, 1, buf.data_ptr()) < 0
    assert L.ria_encode_fixed_frame_batch_dev(h, 4, 1, 0, buf.data_ptr(), 240, 240, 1, buf.data_ptr()) < 0      # bits_per_symbol
    assert L.ria_ofdm_tx_frames_dev(h, C.addressof(cfg), buf.data_ptr(), 324, 324, 1, out.data_ptr(), 100) < 0   # out_stride too small
    assert b"out_stride" in L.ria_last_error(h)
    assert L.ria_burst_deinterleave_batch_dev(h, out.data_ptr(), 2000, 4, 1, out.data_ptr() + 65536, 2592) < 0    # short rows
    assert L.ria_burst_deinterleave_batch_dev(h, out.data_ptr(), 2592, 4, 1, out.data_ptr(), 2592) < 0            # in place
    assert L.ria_ldpc_ladder_perturb_dev(h, out.data_ptr(), 1, 39, out.data_ptr()) < 0                            # attempt range
    mc = mcdpsk.MultiCarrierDPSKConfig.default(1, 4, 10)
    assert L.ria_mcdpsk_tx_frames_dev(h, C.addressof(mc), buf.data_ptr(), 81, 81, 1, out.data_ptr(), 10) < 0
    assert L.ria_mcdpsk_tx_frame_samples(C.addressof(mc), 81) == (8 + 1 + 65 * 4) * 512
    assert L.ria_ofdm_tx_frame_samples(C.addressof(cfg), 324) == 13440
    with pytest.raises(ria_b200.RiaError):
        ctx.set_decode_flags(8)
    assert ctx.get_decode_flags() == 0


def test_device_preambles_are_the_reference_samples(ctx, ref):
    """ria_zc_preamble_dev / ria_chirp_generate_dev (one thread per sample, glibc sinf/cosf incl. the
    large-argument reduction restated in rn_math.h) vs ZCSync::generatePreamble and ChirpSync::generate of
    the unmodified reference: bit-identical."""
    from oracle.bindings import ZcConfig
    from ria_b200 import sync
    zc = ZcConfig.default()
    for frame_type, root in ((0, zc.root_ping), (1, zc.root_pong), (2, zc.root_data), (3, zc.root_control)):
        want = ref.zc_preamble(zc, frame_type)
        got = sync.zc_preamble(sync.ZCConfig.default(), root, ctx=ctx).cpu().numpy()
        assert len(got) == len(want)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), frame_type
    want = ref.chirp_generate()
    got = sync.chirp_generate(ctx=ctx).cpu().numpy()
    assert len(got) == len(want) and np.array_equal(got.view(np.uint32), want.view(np.uint32))

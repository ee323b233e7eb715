"""GPU: the CUDA LDPC decoder through the C ABI, bit-exact against the oracle / golden vectors."""
import os

import numpy as np
import pytest

from oracle.bindings import (R1_4, R1_3, R1_2, R2_3, R3_4, R5_6, RATE_K, RATE_MAX_ITER, awgn_llrs,
                             unpack_bits)

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "ldpc_golden.npz")
RATES = (R1_4, R1_2, R2_3, R3_4, R5_6)


def _decode_gpu(ctx, rate, llr, max_iter, factor, stride=None):
    import torch
    from ria_b200 import fec
    dec = fec.LDPCDecoder(rate, ctx)
    dec.setMaxIterations(max_iter)
    dec.setMinSumFactor(factor)
    info, ok, iters = dec.decode_batch(torch.from_numpy(llr).cuda(), stride)
    torch.cuda.synchronize()
    return info.cpu().numpy(), ok.cpu().numpy(), iters.cpu().numpy()


@pytest.mark.parametrize("rate", RATES)
def test_golden_vectors(ctx, rate):
    g = np.load(GOLD)
    llr = g[f"r{rate}_llr"]
    for tag, factor in (("a", 0.75), ("b", 0.9375)):
        info, ok, iters = _decode_gpu(ctx, rate, llr, RATE_MAX_ITER[rate], factor, 68)
        assert np.array_equal(ok, g[f"r{rate}_{tag}_ok"])
        assert np.array_equal(iters, g[f"r{rate}_{tag}_iters"])
        kb = (RATE_K[rate] + 7) // 8
        assert np.array_equal(info[:, :kb], g[f"r{rate}_{tag}_info"][:, :kb])


@pytest.mark.parametrize("rate", RATES + (R1_3,))
@pytest.mark.parametrize("factor", (0.75, 0.9375))
def test_matches_oracle_on_seeded_awgn(ctx, port, rate, factor):
    """4096 codewords per rate around the waterfall (mix of converging and failing)."""
    rng = np.random.default_rng(7 + rate)
    k = RATE_K[rate]
    n = 4096
    esn0 = {R1_4: -2.2, R1_3: 2.0, R1_2: 2.0, R2_3: 4.4, R3_4: 5.4, R5_6: 6.6}[rate]
    bits = np.zeros((n, 648), np.uint8)
    for i in range(64):                       # 64 distinct codewords, tiled
        cw = port.ldpc_encode(rate, rng.integers(0, 256, size=k // 8, dtype=np.uint8))[:81]
        bits[i::64] = unpack_bits(cw)
    llr = awgn_llrs(bits, esn0, rng)
    want = port.ldpc_decode_batch(rate, llr, RATE_MAX_ITER[rate], factor, 68)
    got = _decode_gpu(ctx, rate, llr, RATE_MAX_ITER[rate], factor, 68)
    assert np.array_equal(got[1], want[1]), "ok flags differ"
    assert np.array_equal(got[2], want[2]), "iteration counts differ"
    assert np.array_equal(got[0], want[0]), "info bytes differ"
    assert 0.02 < want[1].mean() < 0.999     # the case really straddles the waterfall


@pytest.mark.parametrize("max_iter", (0, 1, 2, 7))
def test_iteration_limits(ctx, port, max_iter):
    rng = np.random.default_rng(3)
    rate = R1_2
    cw = port.ldpc_encode(rate, rng.integers(0, 256, size=40, dtype=np.uint8))[:81]
    llr = awgn_llrs(np.tile(unpack_bits(cw), (512, 1)), 2.5, rng)
    want = port.ldpc_decode_batch(rate, llr, max_iter, 0.9375)
    got = _decode_gpu(ctx, rate, llr, max_iter, 0.9375, 64)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)


def test_ragged_batch_sizes_and_empty(ctx, port):
    import torch
    from ria_b200 import fec
    rng = np.random.default_rng(11)
    rate = R3_4
    for n in (1, 7, 8, 9, 255, 1185):
        llr = (rng.standard_normal((n, 648)) * 4 + 3).astype(np.float32)
        want = port.ldpc_decode_batch(rate, llr, 60, 0.9375, 61)
        got = _decode_gpu(ctx, rate, llr, 60, 0.9375, 61)
        for a, b in zip(got, want):
            assert np.array_equal(a, b)
    dec = fec.LDPCDecoder(rate, ctx)
    info, ok, iters = dec.decode_batch(torch.zeros((0, 648), device="cuda"))
    assert info.shape[0] == 0 and ok.numel() == 0 and iters.numel() == 0


def test_host_entry_point_and_reference_semantics(ctx, port, ref):
    """decodeSoft batch=1 semantics incl. multi-block bit concatenation and the short-input pad
    (ldpc_decoder.cpp:284-429) through the host-buffer C ABI entry."""
    from ria_b200 import fec
    rng = np.random.default_rng(5)
    for rate in (R3_4, R1_2, R1_4):
        k = RATE_K[rate]
        dec = fec.LDPCDecoder(rate, ctx)
        dec.setMaxIterations(RATE_MAX_ITER[rate])
        dec.setMinSumFactor(0.9375)
        cws = [port.ldpc_encode(rate, rng.integers(0, 256, size=k // 8, dtype=np.uint8))[:81]
               for _ in range(4)]
        llr = awgn_llrs(np.concatenate([unpack_bits(c) for c in cws]), 8.0 if rate != R1_4 else 1.0, rng)
        for n_llr in (648, 600, 4 * 648, 3 * 648 + 100, 1):
            got = dec.decodeSoft(llr[:n_llr])
            want, ok, it = ref.ldpc_decode_soft(rate, llr[:n_llr], RATE_MAX_ITER[rate], 0.9375)
            assert got == want.tobytes()
            assert dec.lastDecodeSuccess() == ok
            assert dec.lastIterations() == it
        assert dec.decodeSoft(np.zeros(0, np.float32)) == b"" and not dec.lastDecodeSuccess()
        if rate == R3_4:
            codec = fec.LDPCCodec(R3_4, ctx)        # ICodec wrapper: factor 0.75, 60 iterations
            res = codec.decodeExtended(llr[:648])
            want, ok, it = ref.ldpc_decode_soft(rate, llr[:648], 60, 0.75)
            assert res.success == ok and res.iterations == it and res.data == want.tobytes()
            assert codec.getInfoBits() == 486 and codec.getDataBytes() == 60 and codec.getMaxIterations() == 60


def test_full_size_properties(ctx, port):
    """BASELINE config 2 size (1M codewords): size-independent properties -- every codeword the
    decoder reports ok satisfies H.c = 0 on re-encode, noiseless inputs converge in 0 iterations,
    and a 4096-row sample is bit-exact against the oracle."""
    import torch
    from ria_b200 import fec
    rate = R3_4
    k = RATE_K[rate]
    n = 1 << 20
    rng = np.random.default_rng(2026)
    base = np.zeros((256, 648), np.uint8)
    datas = rng.integers(0, 256, size=(256, 61), dtype=np.uint8)
    datas[:, 60] &= 0xFC                          # k = 486 bits: last byte carries 6 bits
    for i in range(256):
        base[i] = unpack_bits(port.ldpc_encode(rate, datas[i])[:81])
    gen = torch.Generator(device="cuda").manual_seed(1)
    s = (1.0 - 2.0 * torch.from_numpy(base).cuda().float()).repeat(n // 256, 1)
    snr = 10 ** (7.0 / 10)
    llr = 2.0 * (s + torch.randn(s.shape, device="cuda", generator=gen) / snr ** 0.5) * snr
    dec = fec.LDPCDecoder(rate, ctx)
    dec.setMaxIterations(60)
    dec.setMinSumFactor(0.9375)
    info, ok, iters = dec.decode_batch(llr)
    torch.cuda.synchronize()
    okf = ok.bool()
    assert okf.float().mean().item() > 0.97
    want_info = torch.from_numpy(datas).cuda().repeat(n // 256, 1)
    # The reference's R3/4 H leaves 162 of the 486 info bits with degree 0 (SURVEY.md section 7
    # "Quirks"): parity can pass while those bits carry raw channel errors.  Bits that ARE in
    # some check must be right whenever the decoder says ok (undetected errors are rare).
    row_ptr, edge_var = fec.get_matrix(rate)
    prot = np.zeros(488, np.uint8)
    prot[edge_var[edge_var < k]] = 1
    assert 324 <= prot[:k].sum() < k          # some info bits are in no check at all
    mask = torch.from_numpy(np.packbits(prot)).cuda()
    wrong = (((info[okf] ^ want_info[okf]) & mask) != 0).any(dim=1).float().mean().item()
    assert wrong < 1e-2                       # weak random code: ~0.2 % undetected at 7 dB
    assert (iters[okf] < 60).all() and (iters[~okf] == 60).all()
    idx = torch.randperm(n, device="cuda", generator=gen)[:4096]
    sub = llr[idx].cpu().numpy()
    w = port.ldpc_decode_batch(rate, sub, 60, 0.9375, 61)
    assert np.array_equal(info[idx].cpu().numpy(), w[0])
    assert np.array_equal(ok[idx].cpu().numpy(), w[1])
    assert np.array_equal(iters[idx].cpu().numpy(), w[2])
    # noiseless: converges at iteration index 0
    info0, ok0, it0 = dec.decode_batch((8.0 * s[:4096]).contiguous())
    assert ok0.all() and (it0 == 0).all() and (info0 == want_info[:4096]).all()
    # padding bytes of a wider stride are zeroed
    info1, _, _ = dec.decode_batch((8.0 * s[:64]).contiguous(), 72)
    assert (info1[:, 61:] == 0).all() and (info1[:, :61] == want_info[:64]).all()

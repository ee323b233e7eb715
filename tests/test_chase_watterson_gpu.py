"""GPU: HARQ chase combining (bit-exact vs fec::ChaseCache) and the Watterson channel
(statistical parity vs sim::WattersonChannel)."""
import numpy as np
import pytest

from oracle.bindings import R1_4, WattersonConfig as RefWatt, awgn_llrs, unpack_bits

pytestmark = pytest.mark.gpu


def test_chase_combine_matches_reference(ctx, ref, port):
    """tools/test_chase_cache.cpp: combining improves decode; accumulation is bit-exact and stops
    after MAX_COMBINES receptions."""
    import torch
    from ria_b200 import fec
    rng = np.random.default_rng(2)
    cache = fec.ChaseCache(ctx=ctx)
    n_keys, n_rx = 12, 6
    cws = [port.ldpc_encode(R1_4, rng.integers(0, 256, size=20, dtype=np.uint8))[:81] for _ in range(n_keys)]
    rx = np.stack([[awgn_llrs(unpack_bits(cw), -5.5, rng) for cw in cws] for _ in range(n_rx)])   # [rx][key][648]
    keys = [(k, 0x111, 0x222) for k in range(n_keys)]
    for r in range(n_rx):
        ok = cache.store_batch(keys, [1] * n_keys, [4] * n_keys, torch.from_numpy(rx[r]).cuda())
        assert all(ok) == (r < 4)                                  # MAX_COMBINES = 4
    torch.cuda.synchronize()
    dec = fec.LDPCDecoder(R1_4, ctx)
    dec.setMaxIterations(50)
    dec.setMinSumFactor(0.9375)
    single_ok = combined_ok = 0
    for k in range(n_keys):
        want, stored, count = ref.chase_combine(rx[:, k], 1, 4)
        got = cache.getCombined(keys[k], 1)
        assert stored == 4 and count == 4 == cache.getCombineCount(keys[k], 1)
        assert np.array_equal(got.cpu().numpy(), want)             # same adds, same order
        _, ok1, _ = dec.decode_batch(torch.from_numpy(rx[0, k:k + 1]).cuda())
        _, ok4, _ = dec.decode_batch(got.reshape(1, 648).contiguous())
        single_ok += int(ok1.item()); combined_ok += int(ok4.item())
    assert combined_ok > single_ok and combined_ok >= n_keys - 1
    for k in range(n_keys):
        cache.markDecoded(keys[k], 1)
        assert cache.getCombined(keys[k], 1) is None
    # LRU eviction at max_entries
    small = fec.ChaseCache(max_entries=2, ctx=ctx)
    x = torch.from_numpy(rx[0, :3]).cuda()
    small.store_batch([("a",), ("b",), ("c",)], [0, 0, 0], [1, 1, 1], x)
    assert small.size() == 2 and small.stats["entries_evicted"] == 1 and small.getCombined(("a",), 0) is None


def test_chase_round_with_more_keys_than_entries_is_sequential(ctx):
    """A round that brings more new keys than max_entries: evictions recycle rows that earlier items of the
    same round wrote, so store_batch must split the launch there.  The result has to be what item-by-item
    ChaseCache::store calls give (chase_cache.cpp:27-88): same surviving entries, same accumulated soft bits."""
    import torch
    from ria_b200 import fec
    rng = np.random.default_rng(5)
    n = 40
    soft = torch.from_numpy(rng.standard_normal((n, 648)).astype(np.float32)).cuda()
    # keys 0..9 twice (second copy combines), then 20 new keys that evict them while the round is open
    keys = [(k % 10, 1, 2) for k in range(20)] + [(100 + k, 1, 2) for k in range(20)]
    cws = [int(c) for c in rng.integers(0, 3, size=n)]
    seen, cw_fixed = set(), []
    for k, c in zip(keys, cws):                      # a (key, cw) pair appears once per round
        while (k, c) in seen:
            c = (c + 1) % 3
        seen.add((k, c)); cw_fixed.append(c)
    batched = fec.ChaseCache(max_entries=8, max_cw=4, ctx=ctx)
    ok_b = batched.store_batch(keys, cw_fixed, [3] * n, soft)
    seq = fec.ChaseCache(max_entries=8, max_cw=4, ctx=ctx)
    ok_s = [seq.store_batch([keys[i]], [cw_fixed[i]], [3], soft[i:i + 1])[0] for i in range(n)]
    torch.cuda.synchronize()
    assert ok_b == ok_s and batched.stats == seq.stats
    assert set(batched.entries) == set(seq.entries) and len(batched.entries) == 8
    for key, e in seq.entries.items():
        for cw in range(3):
            a, b = seq.getCombined(key, cw), batched.getCombined(key, cw)
            assert (a is None) == (b is None)
            if a is not None:
                assert torch.equal(a, b), (key, cw)


def test_watterson_statistics_vs_reference(ctx, ref):
    import torch
    from ria_b200 import sim
    L = 48000
    t = np.arange(L)
    tone = (0.5 * np.sin(2 * np.pi * 1500 * t / 48000)).astype(np.float32)
    pool = torch.from_numpy(tone).cuda().unsqueeze(0)
    for cond, name in ((1, "good"), (3, "poor"), (4, "flutter")):
        cfg = sim.WattersonConfig.preset(cond, 60.0)       # almost noise-free: look at the fading
        n = 384
        out = sim.watterson_batch(cfg, pool, n, seed=3, ctx=ctx)
        again = sim.watterson_batch(cfg, pool, n, seed=3, ctx=ctx)
        assert torch.equal(out, again)
        p_gpu = (out.double() ** 2).mean(dim=1).cpu().numpy() / np.mean(tone.astype(np.float64) ** 2)
        rcfg = RefWatt.from_buffer_copy(bytes(cfg))
        p_ref = []
        for s in range(96):
            # the reference channel starts at f = (1, 0); let it run in before measuring
            y = ref.watterson_process(rcfg, 100 + s, np.tile(tone, 6))[-L:]
            p_ref.append(np.mean(y.astype(np.float64) ** 2) / np.mean(tone.astype(np.float64) ** 2))
        p_ref = np.array(p_ref)
        # mean received power: two paths x 0.707^2 x E|f|^2 (~1) ~ 1, plus interference between the
        # delayed copies; both implementations agree within sampling error
        assert abs(p_gpu.mean() - p_ref.mean()) < 0.25 * p_ref.mean(), (name, p_gpu.mean(), p_ref.mean())
        if cond != 4:      # slow fading: frame powers vary a lot between frames in both
            assert p_gpu.std() > 0.15 * p_gpu.mean() and p_ref.std() > 0.15 * p_ref.mean()
            assert abs(p_gpu.std() / p_gpu.mean() - p_ref.std() / p_ref.mean()) < 0.2
    # noise level: AWGN preset at 10 dB on a constant-envelope tone
    cfg = sim.WattersonConfig.preset(0, 10.0)
    out = sim.watterson_batch(cfg, pool, 64, seed=1, ctx=ctx)
    noise = out - pool
    snr = 10 * np.log10(np.mean(tone.astype(np.float64) ** 2) / (noise.double() ** 2).mean().item())
    # noise_std = rms(NON-SILENT samples) * 10^(-snr/20) (hf_channel.hpp:111-123): the tone has 2 exact
    # zeros per 32-sample period, so the effective SNR is 10 dB - 10 log10(32/30)
    assert abs(snr - (10.0 - 10 * np.log10(32 / 30))) < 0.05
    # multipath only: output = g1 s[n] + g2 s[n - D] exactly
    cfg = sim.WattersonConfig.preset(2, 100.0)
    cfg.fading_enabled = 0
    cfg.noise_enabled = 0
    out = sim.watterson_batch(cfg, pool, 1, ctx=ctx)[0].cpu().numpy()
    D = 48
    want = tone * np.float32(0.707) + np.concatenate([np.zeros(D, np.float32), tone[:-D]]) * np.float32(0.707)
    assert np.allclose(out, want, atol=1e-6)

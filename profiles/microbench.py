#!/usr/bin/env python
"""Stand-alone timing of every batched entry point of SURVEY.md section 8 on synthetic inputs.

    python profiles/microbench.py > profiles/r1_microbench.txt

Each row: the C-ABI entry, the batch, device time per call (CUDA events on the launching stream,
median of the timed repetitions after warm-up), units per second, ALGORITHMIC bytes per unit
(inputs read once + outputs written once) and the resulting fraction of
the measured HBM peak.  Inputs are larger than the 126 MB L2 wherever a useful batch fits in memory.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def peak_gbs():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    try:
        d = json.load(open(p))
        for k in ("hbm_gbs", "hbm_gbps", "copy_gbs"):
            if k in d:
                return float(d[k])
    except Exception:
        pass
    return 6555.5


def timed(fn, reps=7, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    import torch
    import ria_b200
    from ria_b200 import fec, mcdpsk, ofdm, sim, sync, txsynth
    dev = torch.device("cuda", 0)
    ctx = ria_b200.Context(0)
    peak = peak_gbs()
    rows = []

    def add(name, n, ms, bytes_per_unit, unit="frames"):
        gbs = n * bytes_per_unit / (ms * 1e-3) / 1e9
        rows.append((name, n, ms, n / (ms * 1e-3), bytes_per_unit, gbs, gbs / peak, unit))

    with torch.cuda.stream(torch.cuda.Stream(dev)):
        # ---- OFDM presynced demodulator, every modulation family ----
        for mod, rate, label in ((ofdm.QAM64, 4, "QAM64 R3/4 15 pilots"), (ofdm.QAM16, 3, "QAM16 R2/3"),
                                 (ofdm.DQPSK, 2, "DQPSK R1/2 (configs[0])"), (ofdm.QPSK, 2, "QPSK R1/2")):
            cfg = ofdm.ModemConfig.high_throughput(mod) if mod == ofdm.QAM64 else ofdm.ModemConfig.for_waveform(mod, rate)
            pool, _ = txsynth.make_frame_pool(cfg, rate, 8, seed=3)
            n = 131072
            rx = sim.awgn_batch(torch.from_numpy(pool).to(dev), n, 25.0, seed=1, ctx=ctx)
            dem = ofdm.OFDMDemodulator(cfg, ctx)
            out = dem.process_presynced_batch(rx)
            n_llr = int(out["n_llr"][0])
            ms = timed(lambda: dem.process_presynced_batch(rx))
            add(f"ria_ofdm_presynced_batch_dev  {label}", n, ms, rx.shape[1] * 4 + n_llr * 4)
            cfo = torch.full((n,), 3.7, device=dev)
            ph = torch.zeros(n, device=dev)
            ms = timed(lambda: dem.process_presynced_batch(rx, cfo, ph))
            add(f"  same, CFO correction active (+3.7 Hz)", n, ms, rx.shape[1] * 4 + n_llr * 4)
            if mod == ofdm.QAM64:
                w = torch.from_numpy(pool).to(dev)
                lead = torch.zeros((8, 3000), device=dev)
                rows_sync = sim.awgn_batch(torch.cat([lead, w], dim=1).contiguous(), 32768, 25.0, seed=2, ctx=ctx)
                win = rows_sync[:, :8192].contiguous()
                ms = timed(lambda: sync.ofdm_data_sync_batch(cfg, win, None, 0.3, ctx))
                add("ria_ofdm_data_sync_batch_dev  window 8192", 32768, ms, 8192 * 4 + 32, "windows")
                del rows_sync, win
                # OFDM_COX acquisition: [2000 quiet][silent symbol][4 STS][2 LTS][data ...] in a 24000-sample window
                coded = torch.randint(0, 256, (8, 324), dtype=torch.uint8, device=dev)
                cox = ofdm.ofdm_cox_tx_frames(cfg, coded, ctx)
                rows_cox = sim.awgn_batch(torch.cat([torch.zeros((8, 2000), device=dev), cox], dim=1)[:, :24000].contiguous(),
                                          8192, 20.0, seed=4, ctx=ctx)
                res = sync.results(sync.ofdm_cox_search_sync_batch(cfg, rows_cox, 0.8, None, ctx))
                ms = timed(lambda: sync.ofdm_cox_search_sync_batch(cfg, rows_cox, 0.8, None, ctx))
                add(f"ria_ofdm_cox_search_sync_batch_dev  window 24000 ({int(res['detected'].sum())}/8192 found)", 8192, ms,
                    24000 * 4 + 32, "windows")
                del rows_cox, cox
            del rx, out
            torch.cuda.empty_cache()

        # ---- LDPC, decoder only, per rate at its operating point ----
        for rate, esn0 in ((0, 1.0), (2, 4.0), (3, 6.0), (4, 7.0)):
            k = fec.code_params(rate)[0]
            rng = np.random.default_rng(rate)
            cws = np.stack([txsynth.ldpc_encode_bits(rng.integers(0, 2, size=k, dtype=np.uint8), rate) for _ in range(64)])
            n = 1 << 20
            s = (1.0 - 2.0 * torch.from_numpy(cws).to(dev).float()).repeat(n // 64, 1)
            snr = 10 ** (esn0 / 10)
            llr = (2.0 * (s + torch.randn(s.shape, device=dev) / snr ** 0.5) * snr).contiguous()
            dec = fec.LDPCDecoder(rate, ctx)
            dec.setMaxIterations(fec.recommended_iterations(rate))
            dec.setMinSumFactor(0.9375)
            info, ok, iters = dec.decode_batch(llr)
            ms = timed(lambda: dec.decode_batch(llr))
            add(f"ria_ldpc_decode_batch_dev  rate {rate} Es/N0 {esn0} dB (mean {iters.float().mean().item():.1f} it, "
                f"{100 * ok.float().mean().item():.1f}% ok)", n, ms, 648 * 4 + (k + 7) // 8 + 5, "codewords")
            del s, llr
            torch.cuda.empty_cache()

        # ---- complete decodeFixedFrame (first pass / + retry ladder / + false-positive repair) on degraded frames ----
        for mod, rate, snr, label in ((ofdm.QPSK, 2, -9.0, "QPSK R1/2 AWGN -9 dB"), (ofdm.QAM64, 4, 18.5, "QAM64 R3/4 AWGN 18.5 dB")):
            cfg = ofdm.ModemConfig.high_throughput(mod) if mod == ofdm.QAM64 else ofdm.ModemConfig.for_waveform(mod, rate)
            pool, _ = txsynth.make_frame_pool(cfg, rate, 8, seed=5)
            n = 16384
            rx = sim.awgn_batch(torch.from_numpy(pool).to(dev), n, snr, seed=2, ctx=ctx)
            out = ofdm.OFDMDemodulator(cfg, ctx).process_presynced_batch(rx)
            soft = out["llr"]
            bps = cfg.getDataCarriers() * ofdm.getBitsPerSymbol(mod)
            for flags, what in ((0, "first pass"), (1, "+ retry ladder"), (3, "+ ladder + false-positive repair (RIA_DECODE_FULL)")):
                ctx.set_decode_flags(flags)
                _, st = ofdm.decode_fixed_frame_batch(soft, rate, True, bps, ctx)
                ms = timed(lambda: ofdm.decode_fixed_frame_batch(soft, rate, True, bps, ctx), reps=3, warm=1)
                sa = ofdm.status_array(st)
                good = int((sa["all_ok"] & sa["header_valid"] & sa["frame_crc_ok"]).sum())
                add(f"ria_frame_decode_batch_dev  {label}, {what}: {good}/{n} frames valid, "
                    f"{int((sa['ladder_cw_mask'] != 0).sum())} by ladder, {int((sa['fp_repair'] == 1).sum())} repaired", n, ms, 2592 * 4 + 4 * 61 + 40)
            ctx.set_decode_flags(0)
            del rx, out, soft
            torch.cuda.empty_cache()

        # ---- MC-DPSK demodulator ----
        for bits, spread, label in ((1, 4, "DBPSK x4 (configs[2])"), (1, 1, "DBPSK x1"), (2, 1, "DQPSK x1")):
            cfg = mcdpsk.MultiCarrierDPSKConfig.default(bits, spread, 10)
            body = np.stack([txsynth.mcdpsk_modulate_frame(cfg, bytes(np.random.default_rng(i).integers(0, 256, 81, dtype=np.uint8)))
                             for i in range(4)])
            n = 16384 if spread == 4 else 65536
            rx = sim.awgn_batch(torch.from_numpy(body).to(dev), n, 0.0, seed=4, ctx=ctx)
            dem = mcdpsk.MCDPSKDemodulator(cfg, ctx)
            ms = timed(lambda: dem.process_batch(rx))
            add(f"ria_mcdpsk_process_batch_dev  {label}", n, ms, rx.shape[1] * 4 + 652 * 4)
            cfo = torch.full((n,), 12.5, device=dev)
            ms = timed(lambda: dem.process_batch(rx, cfo))
            add("  same, Hilbert CFO correction active (+12.5 Hz)", n, ms, 3 * rx.shape[1] * 4 + 652 * 4)
            del rx
            torch.cuda.empty_cache()

        # ---- synchronisers ----
        zc = sync.ZCSync(None, ctx)
        pre = txsynth.zc_preamble(5)
        row = np.concatenate([np.zeros(1200, np.float32), pre, np.zeros(7120 - 1200 - len(pre), np.float32)])
        n = 32768
        win = sim.awgn_batch(torch.from_numpy(np.stack([row] * 4)).to(dev), n, 0.0, seed=5, ctx=ctx)
        ms = timed(lambda: zc.detect_batch(win, 0.3, sync.ZC_ROOT_MASK_DATA | sync.ZC_ROOT_MASK_CONTROL))
        add("ria_zc_detect_batch_dev  window 7120, roots DATA+CONTROL", n, ms, 7120 * 4 + 32, "windows")
        del win
        # the production search window of StreamingDecoder (streaming_decoder.cpp:423-435): 31 120 samples
        row = np.concatenate([np.zeros(9000, np.float32), pre, np.zeros(31120 - 9000 - len(pre), np.float32)])
        n = 16384
        win = sim.awgn_batch(torch.from_numpy(np.stack([row] * 4)).to(dev), n, 0.0, seed=7, ctx=ctx)
        ms = timed(lambda: zc.detect_batch(win, 0.3, sync.ZC_ROOT_MASK_DATA | sync.ZC_ROOT_MASK_CONTROL))
        add("ria_zc_detect_batch_dev  window 31120 (production), roots DATA+CONTROL", n, ms, 31120 * 4 + 32, "windows")
        del win
        ch = sync.ChirpSync(None, ctx)
        pre = txsynth.chirp_preamble()
        row = np.concatenate([np.zeros(2000, np.float32), pre, np.zeros(120000 - 2000 - len(pre), np.float32)])
        n = 8192
        win = sim.awgn_batch(torch.from_numpy(np.stack([row] * 2)).to(dev), n, -5.0, seed=6, ctx=ctx)
        ms = timed(lambda: ch.detect_dual_batch(win, 0.15, 1024))
        add("ria_chirp_detect_dual_batch_dev  window 120000", n, ms, 120000 * 4 + 32, "windows")
        del win
        torch.cuda.empty_cache()

        # ---- channel simulation and chase combining ----
        cfg = ofdm.ModemConfig.high_throughput(ofdm.QAM64)
        pool, _ = txsynth.make_frame_pool(cfg, 4, 8, seed=3)
        pool_d = torch.from_numpy(pool).to(dev)
        n = 262144
        out = torch.empty((n, pool.shape[1]), device=dev)
        ms = timed(lambda: sim.awgn_batch(pool_d, n, 20.0, seed=8, out=out, ctx=ctx))
        add("ria_channel_awgn_batch_dev  OFDM frames", n, ms, pool.shape[1] * 4)
        wc = sim.WattersonConfig.preset(sim.WattersonConfig.MODERATE, 20.0)
        ms = timed(lambda: sim.watterson_batch(wc, pool_d, n, None, seed=9, out=out, ctx=ctx))
        add("ria_channel_watterson_batch_dev  moderate (1 ms, 1 Hz)", n, ms, pool.shape[1] * 4)
        del out
        n = 1 << 20
        acc = torch.zeros((n, 648), device=dev)
        llr = torch.randn((n, 648), device=dev)
        slot = torch.arange(n, dtype=torch.int32, device=dev)
        first = torch.zeros(n, dtype=torch.uint8, device=dev)
        import ctypes as C
        from ria_b200._lib import lib

        def chase():
            ctx.set_stream(torch.cuda.current_stream(dev))
            ctx.check(lib().ria_chase_combine_batch_dev(ctx.handle, C.c_void_p(acc.data_ptr()), C.c_void_p(slot.data_ptr()),
                                                        C.c_void_p(first.data_ptr()), C.c_void_p(llr.data_ptr()), 648, n))
        ms = timed(chase)
        add("ria_chase_combine_batch_dev  accumulate", n, ms, 3 * 648 * 4, "codewords")

    print(f"# {torch.cuda.get_device_name(0)}; HBM peak used for the last column: {peak:.1f} GB/s")
    print(f"# {'entry point / case':78s} {'batch':>8s} {'ms/call':>9s} {'units/s':>12s} {'B/unit':>8s} {'GB/s':>8s} {'of peak':>8s}")
    for name, n, ms, ups, b, gbs, frac, unit in rows:
        print(f"{name:80s} {n:8d} {ms:9.3f} {ups:12.4g} {b:8d} {gbs:8.1f} {100 * frac:7.1f}%  {unit}")
    ctx.close()


if __name__ == "__main__":
    main()

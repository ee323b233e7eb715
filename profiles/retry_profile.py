#!/usr/bin/env python
"""Degraded OFDM frames through the complete decodeFixedFrame (for ncu captures of the retry / repair kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ria_b200
from ria_b200 import ofdm, sim, txsynth
dev = torch.device('cuda', 0); ctx = ria_b200.Context(0)
for mod, rate, snr in ((ofdm.QAM64, 4, 18.5), (ofdm.QPSK, 2, -7.0)):
    cfg = ofdm.ModemConfig.high_throughput(mod) if mod == ofdm.QAM64 else ofdm.ModemConfig.for_waveform(mod, rate)
    pool, _ = txsynth.make_frame_pool(cfg, rate, 8, seed=5)
    n = 4096
    rx = sim.awgn_batch(torch.from_numpy(pool).to(dev), n, snr, seed=2, ctx=ctx)
    soft = ofdm.OFDMDemodulator(cfg, ctx).process_presynced_batch(rx)["llr"]
    bps = cfg.getDataCarriers() * ofdm.getBitsPerSymbol(mod)
    for _ in range(2):
        _, st = ofdm.decode_fixed_frame_batch(soft, rate, True, bps, ctx, retry_ladder=True, fp_repair=True)
    torch.cuda.synchronize()
    sa = ofdm.status_array(st)
    print(mod, snr, 'valid', int((sa['all_ok'] & sa['header_valid'] & sa['frame_crc_ok']).sum()), 'ladder', int((sa['ladder_cw_mask'] != 0).sum()),
          'repaired', int((sa['fp_repair'] == 1).sum()), 'given up', int((sa['fp_repair'] == 2).sum()))

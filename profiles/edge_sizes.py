#!/usr/bin/env python
"""Small, odd-sized batches through the kernels added in round 2 (OFDM_COX acquisition + transmitter, the rewritten OFDM
data sync, PING energy, the two-frames-per-warp carrier kernel, the ZC-acquired MC-DPSK chain): window lengths at and
around every size limit of those kernels.  A smoke run for edge sizes (the parity tests cover the values)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ria_b200
from ria_b200 import mcdpsk, ofdm, sim, stream, sync, txsynth

ctx = ria_b200.Context(0)
dev = torch.device("cuda", 0)
cfg = ofdm.ModemConfig.high_throughput(ofdm.QAM64)
coded = torch.randint(0, 256, (8, 324), dtype=torch.uint8, device=dev)
# OFDM_COX: transmitter, search on windows of several lengths (incl. too short), correlation tap
cox = ofdm.ofdm_cox_tx_frames(cfg, coded, ctx)
for window in (24000, 9000, 8000, 65536):
    rows = torch.zeros((8, window), device=dev)
    n = min(window - 500, cox.shape[1])
    rows[:, 500:500 + n] = cox[:, :n]
    rows = sim.awgn_batch(rows, 24, 20.0, seed=4, ctx=ctx)
    nf = torch.zeros(24, device=dev)
    r = sync.results(sync.ofdm_cox_search_sync_batch(cfg, rows, 0.8, nf, ctx))
    print("cox window", window, "found", int(r["detected"].sum()))
sync.ofdm_cox_correlation_batch(cfg, rows, torch.arange(24, dtype=torch.int32, device=dev) * 997, ctx)
# data sync: in-noise windows (one pass) and windows that start inside a burst (two passes), odd lengths
pool, _ = txsynth.make_frame_pool(cfg, 4, 8, seed=3)
w = torch.from_numpy(pool).to(dev)
lead = sim.awgn_batch(torch.cat([torch.zeros((8, 3000), device=dev), w], dim=1).contiguous(), 16, 25.0, seed=2, ctx=ctx)
for win in (8192, 10080, 5000, 3400, 13439):
    a = sync.results(sync.ofdm_data_sync_batch(cfg, lead[:, :win].contiguous(), None, 0.3, ctx))
    b = sync.results(sync.ofdm_data_sync_batch(cfg, lead[:, 3000:3000 + win].contiguous(), None, 0.3, ctx))
    print("data sync window", win, "detected", int(a["detected"].sum()), int(b["detected"].sum()))
# PING energy
print("ping", stream.ping_energy_batch(lead[:, :9000].contiguous(), 4608, ctx)["is_ping"].sum())
# carrier kernel (two frames per warp) incl. an odd batch and the CFO path
dem = ofdm.OFDMDemodulator(cfg, ctx)
rx = sim.awgn_batch(w, 37, 25.0, seed=1, ctx=ctx)
dem.process_presynced_batch(rx)
dem.process_presynced_batch(rx, torch.full((37,), 2.5, device=dev), torch.zeros(37, device=dev))
# ZC-acquired MC-DPSK chain
mcfg = mcdpsk.MultiCarrierDPSKConfig.level4_dbpsk(mcdpsk.SPREAD_4X)
cw = torch.randint(0, 256, (4, 81), dtype=torch.uint8, device=dev)
body = mcdpsk.mcdpsk_tx_frames(mcfg, cw, ctx)
pre = sync.zc_preamble(root=2, device=dev, ctx=ctx)
rows = torch.zeros((4, 1000 + pre.numel() + body.shape[1] + 500), device=dev)
rows[:, 1000:1000 + pre.numel()] = pre
rows[:, 1000 + pre.numel():1000 + pre.numel() + body.shape[1]] = body
rows = sim.awgn_batch(rows, 12, 0.0, seed=5, ctx=ctx)
chain = mcdpsk.McdpskRxChain(mcfg, 0, 50, 0.9375, 0.15, ctx)
acc = torch.zeros((12, 648), device=dev)
out = chain.process_batch_zc(rows, body.shape[1], 31120, acc, True)
torch.cuda.synchronize()
print("zc chain ok", int(out["ok"].sum()), "of 12")

#!/usr/bin/env python
"""Per-kernel device time of ria_mcdpsk_process_batch_dev with the Hilbert CFO correction active."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ria_b200
from ria_b200 import mcdpsk, sim, txsynth
dev = torch.device('cuda', 0); ctx = ria_b200.Context(0)
cfg = mcdpsk.MultiCarrierDPSKConfig.default(1, 4, 10)
body = np.stack([txsynth.mcdpsk_modulate_frame(cfg, bytes(np.random.default_rng(i).integers(0, 256, 81, dtype=np.uint8))) for i in range(4)])
n = 16384
rx = sim.awgn_batch(torch.from_numpy(body).to(dev), n, 0.0, seed=4, ctx=ctx)
dem = mcdpsk.MCDPSKDemodulator(cfg, ctx)
for c in (0.0, 12.5):
    cfo = torch.full((n,), c, device=dev) if c else None
    for _ in range(2): dem.process_batch(rx, cfo)
    torch.cuda.synchronize(); ctx.set_timing(True)
    for _ in range(3): dem.process_batch(rx, cfo)
    torch.cuda.synchronize()
    print('cfo', c, {name: round(ctx.get_timing(k)[0] / 3, 3) for k, name in ((9, 'scan+hilbert'), (4, 'demod'))})
    ctx.set_timing(False)

#!/usr/bin/env python
"""One batch of OFDM data-sync windows (the microbench row) for an ncu capture of ofdm_data_sync_kernel."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ria_b200
from ria_b200 import ofdm, sim, sync, txsynth

ctx = ria_b200.Context(0)
dev = torch.device("cuda", 0)
cfg = ofdm.ModemConfig.high_throughput(ofdm.QAM64)
pool, _ = txsynth.make_frame_pool(cfg, 4, 8, seed=3)
w = torch.from_numpy(pool).to(dev)
rows = sim.awgn_batch(torch.cat([torch.zeros((8, 3000), device=dev), w], dim=1).contiguous(), 8192, 25.0, seed=2, ctx=ctx)
win = rows[:, :8192].contiguous()
for _ in range(3):
    out = sync.ofdm_data_sync_batch(cfg, win, None, 0.3, ctx)
torch.cuda.synchronize()
r = sync.results(out)
print("detected", int(r["detected"].sum()), "of", len(r))

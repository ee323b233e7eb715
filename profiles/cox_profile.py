#!/usr/bin/env python
"""One batch of OFDM_COX acquisition windows (the microbench row) for an ncu capture / timing."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ria_b200
from ria_b200 import ofdm, sim, sync

ctx = ria_b200.Context(0)
dev = torch.device("cuda", 0)
cfg = ofdm.ModemConfig.high_throughput(ofdm.QAM64)
n = int(os.environ.get("COX_WINDOWS", "2048"))
coded = torch.randint(0, 256, (8, 324), dtype=torch.uint8, device=dev)
cox = ofdm.ofdm_cox_tx_frames(cfg, coded, ctx)
rows = sim.awgn_batch(torch.cat([torch.zeros((8, 2000), device=dev), cox], dim=1)[:, :24000].contiguous(), n, 20.0, seed=4, ctx=ctx)
for _ in range(2):
    out = sync.ofdm_cox_search_sync_batch(cfg, rows, 0.8, None, ctx)
torch.cuda.synchronize()
print("found", int(sync.results(out)["detected"].sum()), "of", n)
if os.environ.get("COX_TIME"):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        sync.ofdm_cox_search_sync_batch(cfg, rows, 0.8, None, ctx)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("ms per %d windows: %.3f  (%.1f k windows/s)" % (n, ms, n / ms))

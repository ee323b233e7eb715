#!/usr/bin/env python
"""One batch of production-size ZC windows (the microbench row) for an ncu launch list / capture."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ria_b200
from ria_b200 import sim, sync

ctx = ria_b200.Context(0)
dev = torch.device("cuda", 0)
pre = sync.zc_preamble(root=5, device=dev, ctx=ctx)
window = 31120
rows = torch.zeros((8, window), device=dev)
rows[:, 5000:5000 + pre.numel()] = pre
rows = sim.awgn_batch(rows, 4096, 5.0, seed=3, ctx=ctx)
zc = sync.ZCSync(None, ctx)
for _ in range(3):
    out = zc.detect_batch(rows, 0.2, sync.ZC_ROOT_MASK_DATA | sync.ZC_ROOT_MASK_CONTROL, None)
torch.cuda.synchronize()
r = sync.results(out)
print("detected", int(r["detected"].sum()), "of", len(r))
if os.environ.get("ZC_TIME"):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        zc.detect_batch(rows, 0.2, sync.ZC_ROOT_MASK_DATA | sync.ZC_ROOT_MASK_CONTROL, None)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("ms per %d windows: %.3f  (%.3f M windows/s)" % (rows.shape[0], ms, rows.shape[0] / ms / 1e3))

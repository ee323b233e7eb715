#!/usr/bin/env python
"""Device time of ria_chirp_detect_dual_batch_dev for 2048 windows of 120000 samples (env RIA_CHIRP_SUBBATCH)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ria_b200
from ria_b200 import sync
dev = torch.device('cuda', 0); ctx = ria_b200.Context(0)
n, win = 2048, 120000
g = torch.Generator(device=dev); g.manual_seed(1)
rx = torch.randn((n, win), device=dev, generator=g) * 0.1
det = sync.ChirpSync(ctx=ctx) if hasattr(sync, 'ChirpSync') else None
fn = (lambda: det.detect_dual_batch(rx, max_batch=2048)) if det else None
for _ in range(2): fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(torch.cuda.current_stream()); 
for _ in range(3): fn()
e1.record(torch.cuda.current_stream()); torch.cuda.synchronize()
print('sub', os.environ.get('RIA_CHIRP_SUBBATCH'), 'skip_peak', os.environ.get('RIA_CHIRP_SKIP_PEAK'), 'ms per 1024 windows', e0.elapsed_time(e1) / 3 / (n / 1024))

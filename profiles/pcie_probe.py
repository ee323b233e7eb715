#!/usr/bin/env python
"""Raw pinned host -> device bandwidth of the box (what bounds the e2e number)."""
import torch, time
x = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
y = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
for _ in range(2):
    y.copy_(x, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(8):
    y.copy_(x, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"H2D pinned 1 GiB x 8: {8 * (1 << 30) / dt / 1e9:.1f} GB/s")
t0 = time.perf_counter()
for _ in range(8):
    x.copy_(y, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"D2H pinned 1 GiB x 8: {8 * (1 << 30) / dt / 1e9:.1f} GB/s")

#!/usr/bin/env python
"""Quick timing of the GCC-PHAT kernel: pairs/s at config #4 (4096 pairs, L = 24000)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from friture_b200.correlation import GccPhat
P, L = 4096, 24000
d0 = torch.randn(P, L, device="cuda"); d1 = torch.roll(d0, 100, 1) + 0.1 * torch.randn(P, L, device="cuda")
est = GccPhat(L)
for smooth in (False, True):
    est.estimate(d0, d1, smooth=smooth); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): est.estimate(d0, d1, smooth=smooth)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    byts = 2 * L * 4 * (1 + (2 if smooth else 0))
    print("gcc-phat P=%d L=%d smooth=%s: %.3f ms  %.3e pairs/s  %.1f GB/s algorithmic" % (P, L, smooth, ms, P / ms * 1e3, P * byts / ms / 1e6))

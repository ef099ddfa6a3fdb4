#!/usr/bin/env python
"""Throughput of the STFT kernels at the widget's FFT sizes (75 % overlap), not the bench contract."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from friture_b200 import audioproc
p = audioproc()
for n in (256, 1024, 2048, 4096, 8192, 16384):
    hop = n // 4
    C = 64
    F = max(64, (1 << 22) // n)
    x = torch.randn(C, n + (F - 1) * hop, device="cuda") * 0.1
    p.set_fftsize(n)
    out = p.stft(x, hop=hop); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): p.stft(x, hop=hop, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    byts = C * F * (hop * 4 + (n // 2 + 1) * 4)
    print("N=%5d hop=%4d: %8.3f ms  %.3e spectra/s  %.0f GB/s algorithmic" % (n, hop, ms, C * F / ms * 1e3, byts / ms / 1e6), flush=True)

#!/usr/bin/env python
"""Quick STFT timing (config #2 geometry) + parity spot check; not the bench contract."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from friture_b200 import audioproc
from oracle import friture_oracle as fo
from parity import logpower_errors
C, F = 256, 4096
x = torch.randn(C, 2048 + (F - 1) * 1024, device="cuda") * 0.1
out = torch.empty(C, F, 1025, device="cuda")
p = audioproc(); p.set_fftsize(2048)
p.stft(x, hop=1024, out=out); torch.cuda.synchronize()
ref = fo.log_spectrogram(fo.stft_power_batch(x[:4, :2048 + 31 * 1024].cpu().numpy(), 2048, 1024))
print("parity", logpower_errors(out[:4, :32].cpu().numpy(), ref))
for reps in (20, 300):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): p.stft(x, hop=1024, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%s reps=%d: %.3f ms  %.3e spectra/s  %.0f GB/s (%.1f%% of 6572.5)" % (os.environ.get("FRT_B200_LIB", "default"), reps, ms, C * F / ms * 1e3, C * F * 8196 / ms / 1e6, C * F * 8196 / ms / 1e6 / 65.725))

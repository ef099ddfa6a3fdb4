#!/usr/bin/env python
"""Small run of every kernel for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from friture_b200 import audioproc
from friture_b200.octavefilters import Octave_Filters
from friture_b200.correlation import GccPhat

torch.manual_seed(0)
p = audioproc()
for n in (2048, 1024, 64):
    p.set_fftsize(n)
    x = torch.randn(3, n + 5 * (n // 2), device="cuda") * 0.1
    p.stft(x, hop=n // 2, log=True)
    p.stft(x[:, 1:], hop=n // 2, log=False)
    p.stft_host(x.cpu().numpy(), hop=n // 2)
for warps in ("1", "3"):
    os.environ["FRT_BANK_WARPS"] = warps
    for block in (256, 512, 1024):
        bank = Octave_Filters(3)
        x = torch.randn(3, 2 * block, device="cuda") * 0.1
        bank.filter_batch(x, block=block, want_y=True, db=True)
bank = Octave_Filters(24, n_octaves=10)
bank.energies_batch(torch.randn(2, 1024, device="cuda"), block=512)
g = GccPhat(6000)
d0 = torch.randn(3, 6000, device="cuda")
g.estimate(d0, torch.roll(d0, 5, 1), smooth=True, want_xcorr=True)
g.estimate(d0, torch.roll(d0, 5, 1), smooth=True)
torch.cuda.synchronize()
print("sanitize_smoke done")

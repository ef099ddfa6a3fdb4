#!/usr/bin/env python
"""Quick timing of the filterbank kernel (not the bench contract): blocks/s for a few shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from friture_b200.octavefilters import Octave_Filters  # noqa: E402


def run(C, block, nblocks, bpo=3, noct=9, reps=5):
    x = torch.randn(C, block * nblocks, device="cuda") * 0.1
    bank = Octave_Filters(bpo, n_octaves=noct)
    bank.energies_batch(x, block=block)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        bank.energies_batch(x, block=block)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    bps = C * nblocks / (ms * 1e-3)
    print("C=%5d block=%4d nblocks=%4d bpo=%2d noct=%2d: %8.3f ms  %.3e blocks/s  %.3e samples/s"
          % (C, block, nblocks, bpo, noct, ms, bps, bps * block), flush=True)


if __name__ == "__main__":
    run(1024, 512, 256)
    run(1024, 1024, 128)
    run(1024, 256, 512)
    run(1024, 512, 1)
    run(8192, 512, 64)
    run(256, 512, 256)
    run(1024, 512, 256, noct=10)
    run(1024, 512, 64, bpo=24)

"""One launch of the STFT kernel per requested FFT size (for `ncu -k regex:stft_`), 75 % overlap:
python tools/ncu_stft.py [channels] [samples] size [size ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from friture_b200 import audioproc  # noqa: E402

C, T = int(sys.argv[1]), int(sys.argv[2])
x = torch.randn((C, T), dtype=torch.float32, device="cuda") * 0.1
for n in (int(v) for v in sys.argv[3:]):
    p = audioproc()
    p.set_fftsize(n)
    p.stft(x, hop=n // 4)
    torch.cuda.synchronize()

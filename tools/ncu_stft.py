"""One launch of the STFT kernel per requested FFT size (for `ncu -k regex:stft_`), 75 % overlap:
python tools/ncu_stft.py channels samples [hop=H] size [size ...]   (default hop: size/4)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from friture_b200 import audioproc  # noqa: E402

C, T = int(sys.argv[1]), int(sys.argv[2])
rest = sys.argv[3:]
hop = None
if rest and rest[0].startswith("hop="):
    hop = int(rest[0][4:])
    rest = rest[1:]
x = torch.randn((C, T), dtype=torch.float32, device="cuda") * 0.1
for n in (int(v) for v in rest):
    p = audioproc()
    p.set_fftsize(n)
    p.stft(x, hop=hop or n // 4)
    torch.cuda.synchronize()

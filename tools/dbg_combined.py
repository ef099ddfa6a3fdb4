import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from friture_b200.octavefilters import Octave_Filters
from friture_b200 import audioproc
from friture_b200.analyzer import ChannelAnalyzer
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
C = 1024
for T in (131072, 132096, 129 * 1024 + 0):
    x = torch.randn(C, T, device="cuda") * 0.1
    bank = Octave_Filters(3, n_octaves=10)
    nb = T // 1024
    print("bank alone T=%d: %.3f ms" % (T, t(lambda: bank.energies_batch(x, block=1024, db=True))))
an = ChannelAnalyzer(C)
x = torch.randn(C, 129 * 1024, device="cuda") * 0.1
spec = torch.empty(C, 128, 1025, device="cuda"); bands = torch.empty(C, 129, 30, device="cuda")
print("combined overlap: %.3f ms" % t(lambda: an.process(x, spec, bands)))
print("combined sequential: %.3f ms" % t(lambda: an.process(x, spec, bands, overlap=False)))
print("bank via analyzer only: %.3f ms" % t(lambda: an._bank(x, bands)))
print("stft only: %.3f ms" % t(lambda: an.proc.stft(x, hop=1024, log=True, out=spec)))

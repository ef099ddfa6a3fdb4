"""Host-in / host-out time of the combined call by number of time segments (FRT_COMB_NSEG):
python tools/perf_e2e.py [channels] [hops]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from friture_b200.analyzer import ChannelAnalyzer  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
F = int(sys.argv[2]) if len(sys.argv) > 2 else 128
T = (F + 1) * 1024
x = (torch.randn((C, T)) * 0.1).pin_memory()
spec = torch.empty((C, F, 1025)).pin_memory()
bands = torch.empty((C, F + 1, 30)).pin_memory()
an = ChannelAnalyzer(C, 2048, 1024, 3, 10)
for nseg in (2, 4, 6, 8, 10, 12, 16, 24, 32):
    os.environ["FRT_COMB_NSEG"] = str(nseg)
    an.process_host(x, spec, bands)
    an.process_host(x, spec, bands)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4):
        an.process_host(x, spec, bands)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 4 * 1e3
    print("nseg %2d: %.2f ms  %.3g units/s  %.1f GB/s each way" % (nseg, ms, C * F / ms * 1e3,
                                                                 C * T * 4 / ms / 1e6), flush=True)

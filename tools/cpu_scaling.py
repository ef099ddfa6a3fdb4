import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
for nproc in (1, 8, 32, 64, 128):
    if nproc > (os.cpu_count() or 1): break
    b = bench.CpuBaseline(n_channels=max(256, nproc), frames_per_channel=512, nproc=nproc)
    b.step(); n, dt = b.step()
    nv, dtv = b.step(bench._cpu_stft_step_vectorised); nv, dtv = b.step(bench._cpu_stft_step_vectorised)
    b.close()
    print("nproc %3d: loop %.0f spectra/s (%.0f per proc), vectorised %.0f" % (nproc, n/dt, n/dt/nproc, nv/dtv), flush=True)
try:
    print(open('/sys/fs/cgroup/cpu.max').read())
except Exception as e: print(e)
print(len(os.sched_getaffinity(0)))

#!/usr/bin/env python
"""bench.py -- throughput of the Friture spectral hot path on B200 (one JSON line on stdout).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload stft|bank|combined]
                    [--impl ours|reference]

A "step" is one pass of the hot path over one batch of synthetic 48 kHz float32 audio.  The
default workload is BASELINE.json configs[1]: 256 channels (per GPU), 2048-point STFT, hop 1024
(50 % overlap), log-power spectrogram; every (channel, frame) is one `analyzelive` call of the
reference plus `log_spectrogram`.  `metric` is spectra/sec.

  value      whole-job spectra/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e        same metric through the public API with HOST buffers (pinned): H2D + kernels + D2H
             inside the timed region
  roofline   dominant kernel: algorithmic bytes per launch / its CUDA-event duration, against the
             measured HBM peak of MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (a restatement of the reference's NumPy path) on this host's cores
  --impl reference   times that CPU path alone, all host cores (the reference is pure Python and
             cannot travel to the GPU box; oracle/ is its validated restatement)

Multi-GPU: one process per GPU under torchrun; channels are independent streams, so they are
sharded across ranks with no data-path collective (weak scaling: 256 channels per GPU).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_FFT = 2048
HOP = 1024
NBINS = N_FFT // 2 + 1
STFT_BYTES_PER_SPECTRUM = HOP * 4 + NBINS * 4      # 8196 B: each sample read once, each bin written once


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is None:
            return
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            try:
                self.proc.kill()
            except Exception:
                pass

    def summary(self, t0, t1):
        sm, mx, reasons, power = [], [], set(), []
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for (_, r) in self.rows]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            p = [q.strip() for q in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
                power.append(float(p[6]))
            except ValueError:
                continue
            for name, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": float(max(power)) if power else None}


# ----------------------------------------------------------------------------- CPU baseline
_W = {}


def _cpu_init(seed_base, n_ch_total, n_samples, nproc):
    """Worker initialiser: each worker owns a fixed shard of channels of the synthetic input."""
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    _W["args"] = (seed_base, n_ch_total, n_samples, nproc)
    _W["data"] = {}


def _cpu_shard(idx):
    seed_base, n_ch_total, n_samples, nproc = _W["args"]
    if idx not in _W["data"]:
        lo = idx * n_ch_total // nproc
        hi = (idx + 1) * n_ch_total // nproc
        rng = np.random.default_rng(seed_base + idx)
        _W["data"][idx] = (rng.standard_normal((hi - lo, n_samples)) * 0.1).astype(np.float32)
    return _W["data"][idx]


def _cpu_stft_step(idx):
    """The reference's spectrogram loop on this worker's channels: per frame `analyzelive`
    (friture/spectrogram.py:149-159), then log_spectrogram on the column block (:161)."""
    from oracle import friture_oracle as fo
    x = _cpu_shard(idx)
    n = 0
    acc = 0.0
    for c in range(x.shape[0]):
        sp = fo.stft_power(x[c].astype(np.float64), N_FFT, HOP)
        db = fo.log_spectrogram(sp)
        acc += float(db[0, 0])
        n += sp.shape[0]
    return n, acc


def _cpu_stft_step_vectorised(idx):
    from oracle import friture_oracle as fo
    x = _cpu_shard(idx)
    db = fo.log_spectrogram(fo.stft_power_batch(x, N_FFT, HOP))
    return db.shape[0] * db.shape[1], float(db[0, 0, 0])


def usable_cpus():
    """CPUs this process may actually use: min(affinity, cgroup cpu.max quota)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(np.ceil(int(quota) / int(period)))))
    except Exception:
        pass
    return max(1, n)


def best_cpu_baseline(n_channels, frames):
    """The CPU path with the worker count that gives it the highest throughput on this host
    (one or two workers per usable CPU)."""
    ncpu = usable_cpus()
    best = None
    for workers in sorted({ncpu, min(2 * ncpu, os.cpu_count() or ncpu)}):
        base = CpuBaseline(n_channels=max(n_channels, workers), frames_per_channel=frames,
                           nproc=workers)
        base.step()
        n, dt = base.step()
        if best is None or n / dt > best[1]:
            if best is not None:
                best[0].close()
            best = (base, n / dt)
        else:
            base.close()
    best[0].cores = ncpu
    return best[0]


class CpuBaseline:
    """Times the oracle port of the reference path on the host cores (multiprocessing, one
    worker per logical CPU, channels partitioned evenly)."""

    def __init__(self, n_channels, frames_per_channel, nproc=None):
        import multiprocessing as mp
        self.nproc = nproc or os.cpu_count() or 1
        self.n_channels = max(n_channels, self.nproc)
        self.frames = frames_per_channel
        n_samples = N_FFT + (frames_per_channel - 1) * HOP
        ctx = mp.get_context("fork")
        self.pool = ctx.Pool(self.nproc, initializer=_cpu_init,
                             initargs=(4321, self.n_channels, n_samples, self.nproc))
        self.sample = ("%d channels x %d frames (N=%d, hop=%d), per-frame analyzelive loop + log10, "
                       "%d worker processes" % (self.n_channels, self.frames, N_FFT, HOP, self.nproc))

    def step(self, fn=_cpu_stft_step):
        t0 = time.perf_counter()
        res = self.pool.map(fn, range(self.nproc), chunksize=1)
        dt = time.perf_counter() - t0
        return sum(r[0] for r in res), dt

    def close(self):
        self.pool.close()
        self.pool.join()


def _cpu_bank_c(arg):
    """IIR bank + smoothing through the plain-C restatement (oracle/iir_df2t.c)."""
    seed, nch, nblk = arg
    from oracle import friture_oracle as fo
    from oracle import iir_c
    from friture_b200 import filter_data
    bdec, adec, _ = filter_data.decimator()
    boct, aoct, _ = filter_data.bands(3)
    orc = fo.OctaveSpectrumOracle(bdec, adec, list(boct), list(aoct))
    bank = iir_c.BankC(bdec, adec, list(boct), list(aoct), orc.alphas, n_channels=nch)
    x = (np.random.default_rng(seed).standard_normal((nch, 512 * nblk)) * 0.1).astype(np.float32)
    t0 = time.perf_counter()
    bank.process(x, 512)
    return nch * nblk, time.perf_counter() - t0


def _cpu_gcc(arg):
    seed, npairs = arg
    from oracle import friture_oracle as fo
    rng = np.random.default_rng(seed)
    d0 = rng.standard_normal((npairs, 24000))
    d1 = np.roll(d0, 137, axis=1) + 0.1 * rng.standard_normal((npairs, 24000))
    t0 = time.perf_counter()
    for p in range(npairs):
        xc = fo.generalized_cross_correlation(d0[p], d1[p])
        fo.delay_peak(xc)
    return npairs, time.perf_counter() - t0


def cpu_other_rows(workers):
    """CPU path of the other rows on the usable host cores: the IIR bank (C restatement of the
    reference's recursion, all workers; and the literal pure-Python loop on one block for the
    true reference cost) and GCC-PHAT (NumPy, as the reference)."""
    import multiprocessing as mp
    from oracle import friture_oracle as fo
    from friture_b200 import filter_data
    out = {}
    ctx = mp.get_context("fork")
    with ctx.Pool(workers) as pool:
        r = pool.map(_cpu_bank_c, [(100 + i, 4, 64) for i in range(workers)])
        wall = max(t for _, t in r)
        out["bank_27band_block512_c_port"] = {"blocks_per_s": sum(n for n, _ in r) / wall,
                                              "workers": workers, "kind": "port (C, -O2, no FMA)"}
        r = pool.map(_cpu_gcc, [(200 + i, 8) for i in range(workers)])
        wall = max(t for _, t in r)
        out["gcc_phat_L24000_numpy"] = {"pairs_per_s": sum(n for n, _ in r) / wall, "workers": workers,
                                        "kind": "port (numpy.fft, as the reference)"}
    bdec, adec, _ = filter_data.decimator()
    boct, aoct, _ = filter_data.bands(3)
    zis = fo.bank_filtic(bdec, adec, list(boct), list(aoct))
    x = np.random.default_rng(0).standard_normal(512) * 0.1
    t0 = time.perf_counter()
    y = x
    zi = 0
    for j in range(9):          # literal pure-Python recursion of friture/signal/lfilter.py:131-139
        for i in (2, 1, 0):
            fo.lfilter_df2t_loop(boct[i], aoct[i], y, zis[zi]); zi += 1
        yd, _ = fo.lfilter_df2t_loop(bdec, adec, y, zis[zi]); zi += 1
        y = yd[::2]
    out["bank_27band_block512_python_loop_1core"] = {"blocks_per_s": 1.0 / (time.perf_counter() - t0),
                                                     "kind": "literal reference recursion, 1 core"}
    return out


def run_reference_arm(args, rank, world):
    """--impl reference: the CPU path alone.  Under torchrun only rank 0 works."""
    if rank != 0:
        return
    base = best_cpu_baseline(256, 1024)
    for _ in range(max(args.warmup, 1)):
        base.step()
    tot_n, tot_t = 0, 0.0
    for _ in range(args.steps):
        n, dt = base.step()
        tot_n += n
        tot_t += dt
    base.close()
    value = tot_n / tot_t
    line = {
        "impl": "reference", "metric": "spectra/sec", "value": value, "unit": "spectra/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 1),
        "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": "spectra/s", "cores": base.cores,
                         "workers": base.nproc, "kind": "port", "sample": base.sample},
        "e2e": {"value": value, "unit": "spectra/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "note": "reference is pure Python/NumPy and cannot travel to the GPU box; this is oracle/, "
                "its validated restatement (same numpy.fft calls, same per-frame loop), on all host cores",
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- ours
def other_workloads(args, dev, rank, world, barrier):
    """Config #3 (1/3-octave bank + RMS), config #5's per-GPU unit (STFT column + band vector per
    1024-sample hop) and, at N>1, the final NCCL all-gather of spectrogram columns.  Each is
    timed over a few launches with CUDA events (max over ranks); these are explanatory extras."""
    import torch
    import torch.distributed as dist
    from friture_b200 import audioproc
    from friture_b200.octavefilters import Octave_Filters
    from friture_b200.sharded import allgather_channels

    def timed(fn, reps):
        fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    res = {}
    C = 1024
    g = torch.Generator(device="cpu").manual_seed(99 + rank)
    nblk = 64
    x = (torch.randn((C, 1024 * nblk), generator=g, dtype=torch.float32) * 0.1).to(dev)
    for name, bpo, noct, block in (("bank_27band_block512", 3, 9, 512),
                                   ("bank_30band_block512", 3, 10, 512),
                                   ("bank_27band_block1024", 3, 9, 1024)):
        bank = Octave_Filters(bpo, device=dev.index, n_octaves=noct)
        nb = x.shape[1] // block
        ms = timed(lambda: bank.energies_batch(x, block=block, db=True), 3)
        blocks = C * nb * world
        # algorithmic bytes per block: input + band vector + state read/write
        nsec = 2 * bpo + 6
        byts = block * 4 + noct * bpo * 4 + 2 * (noct * nsec * 2 + noct * bpo) * 4 / nb
        # FMA-class instructions per input sample: 5 per biquad step, 2 for x^2 + smoothing
        fma = sum((nsec * 5 + bpo * 2) / 2 ** j for j in range(noct))
        res[name] = {"channels_per_gpu": C, "blocks_per_channel": nb, "ms": ms,
                     "blocks_per_s": blocks / (ms * 1e-3),
                     "hbm_gbs_algorithmic": blocks / world * byts / (ms * 1e-3) / 1e9,
                     "fp32_useful_tflops": 2 * blocks / world * block * fma / (ms * 1e-3) / 1e12,
                     "note": "compute-bound recursion (FP32 issue / dependency latency), not HBM"}
        del bank
    # combined unit of config #5: per channel-hop one log-power column + one 27-band vector
    proc = audioproc(handle=None)
    proc.set_fftsize(N_FFT)
    bank = Octave_Filters(3, device=dev.index)
    xs = x[:, :N_FFT + (nblk - 2) * HOP]
    outc = torch.empty((C, nblk - 1, NBINS), dtype=torch.float32, device=dev)

    def combined():
        proc.stft(xs, hop=HOP, log=True, out=outc)
        bank.energies_batch(x[:, :1024 * (nblk - 1)], block=1024, db=True)
    ms = timed(combined, 3)
    res["combined_stft_plus_27band"] = {"channels_per_gpu": C, "hops_per_channel": nblk - 1, "ms": ms,
                                        "units_per_s": C * (nblk - 1) * world / (ms * 1e-3)}
    # the widgets' default FFT sizes (spectrogram 4096, spectrum 8192), 75 % overlap as the widgets use
    for n_fft in (4096, 8192):
        pw = audioproc(handle=None)
        pw.set_fftsize(n_fft)
        hopw = n_fft // 4
        nfr = (x.shape[1] - n_fft) // hopw + 1
        outw = torch.empty((256, nfr, n_fft // 2 + 1), dtype=torch.float32, device=dev)
        ms = timed(lambda: pw.stft(x[:256], hop=hopw, log=True, out=outw), 3)
        res["stft_%d_overlap75" % n_fft] = {
            "channels_per_gpu": 256, "frames_per_channel": nfr, "ms": ms,
            "spectra_per_s": 256 * nfr * world / (ms * 1e-3),
            "hbm_gbs_algorithmic": 256 * nfr * (hopw + n_fft // 2 + 1) * 4 / (ms * 1e-3) / 1e9}
        del outw, pw
    # config #4: GCC-PHAT delay estimation, 4096 pairs x L = 24000 (12 kHz-rate signals)
    from friture_b200.correlation import GccPhat
    Pn, Lg = 4096, 24000
    d0 = torch.randn((Pn, Lg), generator=g, dtype=torch.float32).to(dev)
    d1 = torch.roll(d0, 137, 1) + 0.1 * torch.randn((Pn, Lg), generator=g, dtype=torch.float32).to(dev)
    est = GccPhat(Lg)
    ms = timed(lambda: est.estimate(d0, d1, smooth=False), 3)
    idx, _, _ = est.estimate(d0, d1, smooth=False)
    res["gcc_phat_4096pairs_L24000"] = {"ms": ms, "pairs_per_s": Pn * world / (ms * 1e-3),
                                        "hbm_gbs_algorithmic": Pn * 2 * Lg * 4 / (ms * 1e-3) / 1e9,
                                        "delays_recovered": bool((idx == 137).all().item())}
    del d0, d1, est
    if world > 1:
        # final all-gather of the spectrogram columns over NVLink (north_star); link-bound:
        # every GPU must receive (world-1)/world of ALL columns
        full = torch.empty((C * world, nblk - 1, NBINS), dtype=torch.float32, device=dev)

        def stft_and_gather():
            proc.stft(xs, hop=HOP, log=True, out=outc)
            allgather_channels(outc, C * world, out=full)
        ms_g = timed(stft_and_gather, 3)
        ms_s = timed(lambda: proc.stft(xs, hop=HOP, log=True, out=outc), 3)
        recv = outc.numel() * 4 * (world - 1)
        # the same gather after the spectrum widget's per-tick reduction (one smoothed column per
        # channel per tick instead of one per frame): the payload shrinks by the frames per tick
        from friture_b200.spectrum import SpectrumAnalyzer
        an = SpectrumAnalyzer(C, fft_size=N_FFT, overlap=0.5, response_time=0.125)
        full_tick = torch.empty((C * world, NBINS), dtype=torch.float32, device=dev)

        def tick_and_gather():
            db, _, _ = an.process(xs)
            allgather_channels(db, C * world, out=full_tick)
        ms_t = timed(tick_and_gather, 3)
        res["spectrum_tick_with_allgather"] = {"channels_per_gpu": C, "frames_per_tick": nblk - 1,
                                               "ms": ms_t,
                                               "spectra_per_s": C * (nblk - 1) * world / (ms_t * 1e-3)}
        res["stft_with_allgather"] = {"channels_per_gpu": C, "frames": nblk - 1, "ms_stft": ms_s,
                                      "ms_stft_plus_allgather": ms_g,
                                      "spectra_per_s": C * (nblk - 1) * world / (ms_g * 1e-3),
                                      "allgather_recv_gbs_per_gpu": recv / max(ms_g - ms_s, 1e-6) / 1e6}
    return res


def workload_config(args, world):
    return {"workload": "configs[1]: %d ch/GPU x %d frames, 48 kHz, 2048-pt STFT hop 1024 (50%% overlap) "
                        "+ log-power spectrogram" % (args.channels, args.frames),
            "n_fft": N_FFT, "hop": HOP, "channels_per_gpu": args.channels,
            "frames_per_channel": args.frames, "global_channels": args.channels * world,
            "parallelism": "channel-sharded x%d, no data-path collective" % world,
            "l2_policy": "inputs (%.1f GB/GPU) and outputs larger than L2; no flush needed"
                         % (args.channels * (N_FFT + (args.frames - 1) * HOP) * 4 / 1e9)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="stft", choices=["stft"])
    ap.add_argument("--channels", type=int, default=256, help="channels per GPU")
    ap.add_argument("--frames", type=int, default=4096, help="frames per channel per step")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-others", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; friture_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")   # keep NCCL's banner off stdout (one JSON line)
        dist.init_process_group("nccl", device_id=dev)

    from friture_b200 import audioproc
    from friture_b200._lib import default_handle
    from oracle import friture_oracle as fo
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from parity import logpower_errors, logpower_ok

    handle = default_handle(local_rank)
    proc = audioproc(handle)
    proc.set_fftsize(N_FFT)
    C, F = args.channels, args.frames
    T = N_FFT + (F - 1) * HOP
    gen = torch.Generator(device="cpu").manual_seed(1234 + rank)
    # synthetic broadband audio (sigma 0.1), generated in chunks to bound host memory
    x = torch.empty((C, T), dtype=torch.float32, device=dev)
    for c0 in range(0, C, 16):
        c1 = min(C, c0 + 16)
        x[c0:c1] = (torch.randn((c1 - c0, T), generator=gen, dtype=torch.float32) * 0.1).to(dev)
    out = torch.empty((C, F, NBINS), dtype=torch.float32, device=dev)

    # ---- parity gate (oracle as checker) on a slice of the very buffers being timed
    proc.stft(x, hop=HOP, log=True, out=out)
    torch.cuda.synchronize()
    cs, fs = min(C, 4), min(F, 32)
    ref = fo.log_spectrogram(fo.stft_power_batch(x[:cs, :N_FFT + (fs - 1) * HOP].cpu().numpy(), N_FFT, HOP))
    perr = logpower_errors(out[:cs, :fs].cpu().numpy(), ref)
    if not logpower_ok(perr):
        raise SystemExit("bench.py: parity gate failed: %r" % (perr,))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: K steps, one kernel launch per step
    for _ in range(args.warmup):
        proc.stft(x, hop=HOP, log=True, out=out)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    barrier()
    launches0 = handle.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    t_wall0 = time.time()
    ev[0].record()
    for i in range(args.steps):
        proc.stft(x, hop=HOP, log=True, out=out)
        ev[i + 1].record()
    barrier()
    t_wall1 = time.time()
    launches = handle.launch_count - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    time.sleep(0.2)
    sampler.stop()
    clocks = sampler.summary(t_wall0, t_wall1)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    ms_per_step = total_ms_max / args.steps
    spectra_per_step = C * F * world
    value = spectra_per_step / (ms_per_step * 1e-3)

    peak, peak_src = load_peaks()
    kern_ms = float(np.mean(per_launch_ms))
    achieved = C * F * STFT_BYTES_PER_SPECTRUM / (kern_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "stft2048_kernel<LOGPOWER,VEC>", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": C * F * STFT_BYTES_PER_SPECTRUM,
                "bytes_per_spectrum": STFT_BYTES_PER_SPECTRUM, "kernel_ms": kern_ms}
    traffic_path = os.path.join(ROOT, "profiles", "stft_traffic.json")
    if os.path.isfile(traffic_path):
        with open(traffic_path) as f:
            tr = json.load(f)
        if tr.get("channels") == C and tr.get("frames") == F:
            roofline["traffic"] = tr.get("dram_bytes_per_launch")

    # ---- end to end through the public API with pinned host buffers
    e2e = None
    if not args.no_e2e:
        xh = torch.empty((C, T), dtype=torch.float32, pin_memory=True)
        xh.copy_(x)
        oh = torch.empty((C, F, NBINS), dtype=torch.float32, pin_memory=True)
        proc.stft_host(xh, hop=HOP, log=True, out=oh)   # warm-up (allocates staging)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            proc.stft_host(xh, hop=HOP, log=True, out=oh)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": spectra_per_step * args.e2e_steps / dt, "unit": "spectra/s",
               "h2d_bytes_per_step": int(xh.numel() * 4), "d2h_bytes_per_step": int(oh.numel() * 4),
               "steps": args.e2e_steps, "ms_per_step": 1e3 * dt / args.e2e_steps,
               "api": "audioproc.stft_host -> frt_stft_process_host (pinned host in/out)"}
        same = bool(torch.equal(oh[:2, :8], out[:2, :8].cpu()))
        e2e["matches_device_path"] = same
        del xh, oh

    # ---- other rows of the hot path, short runs (extra information, not the headline value)
    others = {}
    if not args.no_others:
        others = other_workloads(args, dev, rank, world, barrier)

    # ---- CPU baseline on this host (rank 0, N=1 only)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        base = best_cpu_baseline(256, 1024)
        n, dt = base.step()
        n2, dt2 = base.step()
        nv, dtv = base.step(_cpu_stft_step_vectorised)
        nv, dtv = base.step(_cpu_stft_step_vectorised)
        base.close()
        cpu_baseline = {"value": (n + n2) / (dt + dt2), "unit": "spectra/s", "cores": base.cores,
                        "workers": base.nproc, "kind": "port", "sample": base.sample,
                        "vectorised_numpy_value": nv / dtv}
        if not args.no_others:
            cpu_baseline["other_rows"] = cpu_other_rows(base.cores)

    if rank == 0:
        line = {
            "metric": "spectra/sec", "value": value, "unit": "spectra/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args, world),
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "clocks": clocks,
            "gpu_launches": int(launches), "parity": perr, "other_workloads": others,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

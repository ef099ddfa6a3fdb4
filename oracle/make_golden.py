#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported in place from
/root/reference via oracle/ref_import.py).  TEST INFRASTRUCTURE ONLY; run in the build container:

    python oracle/make_golden.py

The fixtures pin the oracle restatement (tests/test_oracle_golden.py, runs anywhere) and are the
reference-derived anchor of the GPU parity tests on the GPU box, where /root/reference does not
exist.  Inputs are stored as float32 (what the GPU path consumes; the reference widens to float64,
friture/audiobackend.py:466-468), outputs as float64.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    ref = ref_import.load()
    os.makedirs(OUT, exist_ok=True)

    # (i) audioproc.analyzelive on seeded white noise, N in {1024, 2048, 8192}
    d = {}
    for n_fft in (1024, 2048, 8192):
        rng = np.random.default_rng(1000 + n_fft)
        x = (rng.standard_normal(n_fft) * 0.1).astype(np.float32)
        proc = ref.audioproc.audioproc()
        proc.set_fftsize(n_fft)
        d["x_%d" % n_fft] = x
        d["power_%d" % n_fft] = proc.analyzelive(x.astype(np.float64))
        d["window_%d" % n_fft] = proc.window
        A, B, C = proc.get_freq_weighting()
        d["freq_%d" % n_fft] = proc.get_freq_scale()
        d["A_%d" % n_fft], d["B_%d" % n_fft], d["C_%d" % n_fft] = A, B, C
    np.savez_compressed(os.path.join(OUT, "analyzelive.npz"), **d)

    # (ii) spectrogram framing loop: ring buffer + data_indexed + analyzelive + log10
    #      (friture/spectrogram.py:131-162), 4 channels x 16 hops, N=2048, hop=1024
    n_fft, hop, nch, nhops = 2048, 1024, 4, 16
    rng = np.random.default_rng(2024)
    x = (rng.standard_normal((nch, n_fft + (nhops - 1) * hop)) * 0.1).astype(np.float32)
    proc = ref.audioproc.audioproc()
    proc.set_fftsize(n_fft)
    cols = np.zeros((nch, nhops, n_fft // 2 + 1))
    for c in range(nch):
        rb = ref.ringbuffer.RingBuffer()
        old_index = 0
        k = 0
        xs = x[c].astype(np.float64)
        # feed 512-frame chunks like the audio backend (audiobackend.py:460-461)
        for p in range(0, xs.shape[0], 512):
            rb.push(xs[None, p:p + 512], 0.0)
            if k == 0:
                need = n_fft          # first frame = samples [0, n_fft)
                if rb.offset >= need:
                    old_index = n_fft
                    cols[c, k] = proc.analyzelive(rb.data_indexed(old_index, n_fft)[0, :])
                    k += 1
            while k > 0 and k < nhops and rb.offset - old_index >= hop:
                old_index += hop
                cols[c, k] = proc.analyzelive(rb.data_indexed(old_index, n_fft)[0, :])
                k += 1
        assert k == nhops, k
    np.savez_compressed(os.path.join(OUT, "spectrogram.npz"), x=x, n_fft=n_fft, hop=hop,
                        logpower=10. * np.log10(cols + 1e-30), power=cols)

    # (iii) IIR bank (friture/filter.py:86-118) + widget smoothing (octavespectrum.py:101-121),
    #       bpo=3, 16 blocks of 512, plus the same stream re-blocked at 256 / 1024
    from friture.filter import (octave_filter_bank_decimation, octave_filter_bank_decimation_filtic,
                                NOCTAVE)
    from friture.signal.exp_smoothing import exp_smoothed_value
    P = ref.generated_filters.PARAMS
    bdec, adec = np.array(P["dec"][0]), np.array(P["dec"][1])
    rng = np.random.default_rng(31)
    xs = (rng.standard_normal(16 * 512) * 0.1).astype(np.float32)
    d = {"x": xs}
    for bpo in (3, 12):
        boct = [np.array(f) for f in P[str(bpo)][0]]
        aoct = [np.array(f) for f in P[str(bpo)][1]]
        for block in (256, 512, 1024):
            if bpo != 3 and block != 512:
                continue
            zis = octave_filter_bank_decimation_filtic(bdec, adec, boct, aoct)
            decs = [2 ** j for j in range(0, NOCTAVE)[::-1] for _ in range(bpo)]
            w = 0.65
            alphas = [1. - (1. - w) ** (1. / (1.0 * 48000 / dec + 1)) for dec in decs]
            kernels = [(1. - a) ** np.arange(int(2 * 4096 / dec) - 1, -1, -1)
                       for a, dec in zip(alphas, decs)]
            disp = [0] * (NOCTAVE * bpo)
            E = []
            ys = [[] for _ in range(NOCTAVE * bpo)]
            for b in range(xs.shape[0] // block):
                y, dec, zis = octave_filter_bank_decimation(
                    bdec, adec, boct, aoct, xs[b * block:(b + 1) * block].astype(np.float64), zis)
                disp = [exp_smoothed_value(k, a, yy ** 2, old)
                        for yy, k, a, old in zip(y, kernels, alphas, disp)]
                E.append(np.array(disp))
                for k in range(len(y)):
                    ys[k].append(y[k])
            d["energies_bpo%d_block%d" % (bpo, block)] = np.array(E)
            if block == 512:
                d["dec_bpo%d" % bpo] = np.array(dec)
                for k in (0, NOCTAVE * bpo // 2, NOCTAVE * bpo - 1):
                    d["y_bpo%d_band%d" % (bpo, k)] = np.concatenate(ys[k])
                d["zis_bpo%d" % bpo] = np.concatenate(zis)

    # (iv) the live FFT-OLA bank on the same stream (secondary comparison, ~5e-4)
    of = ref.octavefilters.Octave_Filters(3)
    acc = np.zeros(27)
    for b in range(16):
        y, dec = of.filter(xs[b * 512:(b + 1) * 512].astype(np.float64))
        acc += np.array([np.sum(v ** 2) for v in y])
    d["fft_bank_energy_sum_bpo3"] = acc
    d["fi_bpo3"], d["flow_bpo3"], d["fhigh_bpo3"] = of.fi, of.flow, of.fhigh
    d["A_bpo3"], d["B_bpo3"], d["C_bpo3"] = of.A, of.B, of.C
    d["f_nominal_bpo3"] = np.array(of.f_nominal)
    np.savez_compressed(os.path.join(OUT, "octave_bank.npz"), **d)

    # (iv-b) band outputs of the LIVE path itself (parity target of Octave_Filters(mode="fft"))
    out = {}
    for bpo in (3, 12):
        of = ref.octavefilters.Octave_Filters(bpo)
        ys = [[] for _ in range(NOCTAVE * bpo)]
        for b in range(16):
            y, dec = of.filter(xs[b * 512:(b + 1) * 512].astype(np.float64))
            for k in range(NOCTAVE * bpo):
                ys[k].append(y[k])
        for k in (0, NOCTAVE * bpo // 2, NOCTAVE * bpo - 1):
            out["y_bpo%d_band%d" % (bpo, k)] = np.concatenate(ys[k])
        out["energy_sum_bpo%d" % bpo] = np.array([np.sum(np.concatenate(v) ** 2) for v in ys])
        out["dec_bpo%d" % bpo] = np.array(dec)
    of = ref.octavefilters.Octave_Filters(3)
    ys = [[] for _ in range(27)]
    for b in range(8):
        y, _ = of.filter(xs[b * 1024:(b + 1) * 1024].astype(np.float64))
        for k in range(27):
            ys[k].append(y[k])
    out["energy_sum_bpo3_block1024"] = np.array([np.sum(np.concatenate(v) ** 2) for v in ys])
    np.savez_compressed(os.path.join(OUT, "octave_bank_fft.npz"), **out)

    # (v) GCC-PHAT, L = 24000 with a known integer delay
    L = 24000
    rng = np.random.default_rng(404)
    base = rng.standard_normal(L)
    d0 = base.astype(np.float32)
    d1 = (np.roll(base, 137) + 0.1 * rng.standard_normal(L)).astype(np.float32)
    xc = ref.correlation.generalized_cross_correlation(d0.astype(np.float64), d1.astype(np.float64))
    d = {"d0": d0, "d1": d1, "xcorr": xc, "argmax": int(np.argmax(np.abs(xc)))}
    # second frame for the temporal smoothing (delay_estimator.py:134-139)
    d0b = rng.standard_normal(L).astype(np.float32)
    d1b = (np.roll(d0b.astype(np.float64), 137) + 0.1 * rng.standard_normal(L)).astype(np.float32)
    xcb = ref.correlation.generalized_cross_correlation(d0b.astype(np.float64), d1b.astype(np.float64))
    sm = 0.3 * xcb + 0.7 * xc
    d.update({"d0b": d0b, "d1b": d1b, "xcorr_b": xcb, "smoothed_b": sm,
              "argmax_b": int(np.argmax(np.abs(sm)))})
    # decimate_multiple (delay_estimator.py:97-98): 48 kHz -> 12 kHz
    from friture.signal.decimate import decimate_multiple, decimate_multiple_filtic
    xin = (rng.standard_normal(2048) * 0.1).astype(np.float32)
    zf = decimate_multiple_filtic(2, bdec, adec)
    o1, zf = decimate_multiple(2, bdec, adec, xin[:1024].astype(np.float64), zf)
    o2, zf = decimate_multiple(2, bdec, adec, xin[1024:].astype(np.float64), zf)
    d.update({"dec_in": xin, "dec_out": np.concatenate([o1, o2])})
    np.savez_compressed(os.path.join(OUT, "gcc_phat.npz"), **d)

    # (vi) exp_smoothed_value_2d across frames (spectrum.py:158, setresponsetime :196-218)
    from friture.signal.exp_smoothing import exp_smoothed_value_2d
    rng = np.random.default_rng(55)
    data = rng.random((1025, 7))
    prev = rng.random(1025)
    n = 0.125 * 48000 / 1024.
    alpha = 1. - (1. - 0.65) ** (1. / (n + 1))
    kernel = (1. - alpha) ** np.arange(8192 - 1, -1, -1)
    np.savez_compressed(os.path.join(OUT, "exp_smoothing.npz"), data=data, prev=prev, alpha=alpha,
                        out=exp_smoothed_value_2d(kernel, alpha, data, prev))

    # (vii) display chain: the reference's Frequency_Resampler and Online_Linear_2D_resampler on
    #       three ticks of scaled dB columns (Mel grid, 20 Hz..24 kHz, 96 rows), then the colour LUT
    #       (Color_Transform needs PyQt6's QColor; its .rgb() word is restated as
    #       0xFF000000 | r<<16 | g<<8 | b on the reference's generated_cmrmap.CMAP)
    from fractions import Fraction
    import friture.plotting.frequency_scales as fscales
    from friture.signal.frequency_resampler import Frequency_Resampler
    from friture.signal.online_linear_2D_resampler import Online_Linear_2D_resampler
    from friture.signal.lookup_table import color_from_float_2D
    from friture.plotting import generated_cmrmap
    cmap = generated_cmrmap.CMAP
    lut = np.array([0xFF000000 | (int(c[0] * 255) << 16) | (int(c[1] * 255) << 8) | int(c[2] * 255)
                    for c in cmap], dtype=np.uint32)
    rng = np.random.default_rng(77)
    nb, height = 1025, 96
    freq = np.linspace(0, 24000, nb)
    fr = Frequency_Resampler(fscales.Mel, 20., 24000., height)
    fr.setfreq(freq)
    L, M = Fraction(48000, 2048) / (Fraction(1) - Fraction(3, 4)) / 1000, Fraction(700, 10000)
    tr = Online_Linear_2D_resampler(L, M, height)
    d = {"lut": lut, "xscaled": fr.xscaled, "ratio": float(L) / float(M)}
    for tick, ncols in enumerate((5, 1, 9)):
        db = (-140.0 + 150.0 * rng.random((nb, ncols))).astype(np.float32)     # dB columns
        norm = (db.astype(np.float64) - (-140.0)) / (0.0 - (-140.0))            # spectrogram.py:127-129
        res = tr.push(fr.push(norm))
        d["db_%d" % tick] = db
        d["resampled_%d" % tick] = res
        d["pixels_%d" % tick] = color_from_float_2D(lut, np.clip(res, 0., 1.))
    np.savez_compressed(os.path.join(OUT, "display.npz"), **d)

    # coefficients carried over from the reference (pins friture_b200/data/filters.npz)
    d = {"bdec": bdec, "adec": adec}
    for bpo in (1, 3, 6, 12, 24):
        d["b%d" % bpo] = np.array(P[str(bpo)][0])
        d["a%d" % bpo] = np.array(P[str(bpo)][1])
    np.savez_compressed(os.path.join(OUT, "coefficients.npz"), **d)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
